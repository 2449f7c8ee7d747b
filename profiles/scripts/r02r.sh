timeout 900 python -m pytest tests/test_gpu_image.py -q 2>&1 | tail -12
