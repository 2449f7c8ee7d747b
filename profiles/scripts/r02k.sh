mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gwkl.py -m gpu -x -q > gpurun_out/pytest_gwkl.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/pytest_gwkl.log
timeout 600 python bench_extra.py --only gwdb 2> gpurun_out/gwdb.err | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in ('ms_per_pair','iterations','gw_dist','lmo_stats','n1000','speedup_vs_cpu_port')})
"
tail -3 gpurun_out/gwdb.err
