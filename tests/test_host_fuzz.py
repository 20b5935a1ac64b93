"""The host parse layer (csrc/host: container, hvcC, SPS/PPS, slice headers — the reference's src/heif, src/hevc) must
reject malformed files with an error code, never read out of bounds or overflow: an ASan + UBSan build of it is driven
over seeded mutations of the fixture (tests/fuzz/host_fuzz.cc, test infrastructure only)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "halfmoonbay.heic")


@pytest.fixture(scope="module")
def fuzz_bin():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "fuzz"), "all"])
    return os.path.join(HERE, "fuzz", "_build", "host_fuzz")


@pytest.fixture(scope="module")
def parser_fuzz_bin(fuzz_bin):
    return os.path.join(HERE, "fuzz", "_build", "parser_fuzz")


@pytest.mark.parametrize("seed0", [0, 100000])
def test_mutated_files_are_rejected_or_parsed_without_memory_errors(fuzz_bin, seed0):
    r = subprocess.run([fuzz_bin, FIXTURE, str(seed0), "1000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    assert lines[-1].startswith("done ok=") and lines[-2].startswith("parsers ok=")
    ok, err = (int(x.split("=")[1]) for x in lines[-1].split()[1:])     # mutated files through the file API
    pok, perr = (int(x.split("=")[1]) for x in lines[-2].split()[1:])   # mutated SPS / PPS / slice headers, parsed directly
    # both outcomes occur: the mutations reach the parsers
    assert ok + err == 1000 and ok > 0 and err > 0
    assert pok + perr >= 3000 and pok > 0 and perr > 0


def test_unmutated_file_parses(fuzz_bin):
    r = subprocess.run([fuzz_bin, FIXTURE, "0", "0"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip().endswith("done ok=0 err=0")  # exits 2 if the fixture did not parse


@pytest.mark.parametrize("seed0", [0, 500000])
def test_device_slice_data_parser_on_corrupt_streams_has_no_memory_errors(parser_fuzz_bin, seed0):
    """The CUDA syntax walker (cabac_parse.cuh) in its host build, exact-size heap arenas, under ASan + UBSan: bit flips,
    garbage, truncation, all-ones / all-zeros, moved entry points, extreme slice QPs and coding tools switched on that
    the stream was not coded with.  compute-sanitizer is not available on the GPU pool; this is its stand-in."""
    r = subprocess.run([parser_fuzz_bin, FIXTURE, str(seed0), "800"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-3000:]
    last = r.stdout.strip().splitlines()[-1]
    ok, flagged = (int(x.split("=")[1]) for x in last.split()[1:])
    assert ok + flagged == 800 and ok >= 1 and flagged > 700  # the unmutated first stream parses; corruption is detected
