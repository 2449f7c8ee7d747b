mkdir -p gpurun_out
python profiles/dropin_profile.py 2>&1 | grep -v "^$" | head -30
python -m pytest tests/test_gpu_dropin.py tests/test_gpu_nimagenet.py -x -q 2>&1 | tail -3
