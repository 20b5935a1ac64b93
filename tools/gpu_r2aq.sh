timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 2
timeout 300 python tools/profile_batch.py --decodes 1 --stage-reps 2 2>&1 | tail -n 2
