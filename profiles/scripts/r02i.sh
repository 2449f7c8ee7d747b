mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench_extra.py --only config3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['workload'][:60], round(d['ms_per_step'],4), 'ms frac', round(d['roofline']['frac'],3))
"
