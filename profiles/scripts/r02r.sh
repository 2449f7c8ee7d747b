timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "fused" 2>&1 | tail -5
