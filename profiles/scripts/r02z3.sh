mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mdjit.py -m gpu -x -q 2>&1 | tail -15
PYTHONPATH=$PWD timeout 300 python profiles/generic_md_workload.py 2>&1 | tee gpurun_out/generic_md.jsonl | cut -c1-330
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log
