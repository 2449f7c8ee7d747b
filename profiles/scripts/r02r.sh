timeout 900 python -m pytest tests/test_gpu_dropin.py -q -k "otmi" 2>&1 | tail -25
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_dropin.py -q -k "otmi_prepare and not 50000" 2>&1 | tail -6
