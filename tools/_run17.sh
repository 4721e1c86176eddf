python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=30 -k "n8192 or with_obs_large" --durations=8 2>&1 | tail -20
