"""Regenerates tests/golden/synth_hashes.json: SHA-256 of the planes FFmpeg decodes from every synthetic stream of
tests/synth/configs.py (and of the stream bytes, pinning the generator's determinism).  Authoring container only."""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ffmpeg_oracle import FFmpegHevc
from tests.synth import synth
from tests.synth.configs import CONFIGS, SEEDS

ff = FFmpegHevc()
out = {"source": "FFmpeg libavcodec 62 native hevc decoder on tests/synth streams", "streams": {}}
for name, cfg in CONFIGS:
    for seed in SEEDS:
        pic = synth.encode(seed, **cfg)
        planes = ff.decode_picture(pic.annexb())
        out["streams"][f"{name}/{seed}"] = {"annexb_sha256": hashlib.sha256(pic.annexb()).hexdigest(), "bytes": len(pic.annexb()),
                                            "planes": [hashlib.sha256(p.tobytes()).hexdigest() for p in planes]}
json.dump(out, open(os.path.join(ROOT, "tests/golden/synth_hashes.json"), "w"), indent=1)
print(len(out["streams"]), "streams")
