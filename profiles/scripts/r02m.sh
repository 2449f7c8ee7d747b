mkdir -p gpurun_out
python profiles/auction_workload.py 1000
python profiles/auction_workload.py 300
timeout 600 python -m pytest tests/test_gpu_gwkl.py -m gpu -x -q > gpurun_out/pytest_gwkl.log 2>&1; echo pytest rc=$?
tail -3 gpurun_out/pytest_gwkl.log
