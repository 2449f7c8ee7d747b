# A/B: zone-specialised accumulate (switch on the SBN window mask, straight-line atomics per zone and class)
mkdir -p gpurun_out
L=$PWD/event_representation_study_b200/lib
for rep in 1 2; do
for v in base new; do
  f=$L/libevrep_$v.so; [ $v = new ] && f=$L/libevrep.so
  EVREP_LIB=$f timeout 120 python bench.py --no-cpu --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$v', round(d['value'],2), 'Gev/s', d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', round(d['e2e']['value'],2))"
  EVREP_LIB=$f timeout 120 python bench.py --no-cpu --no-extras --clustered 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$v clustered', round(d['value'],2), 'Gev/s', d['ms_per_step'], d['roofline']['kernel_ms']['tile kernel'])"
done; done | tee gpurun_out/z11.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log
PYTHONPATH=$PWD timeout 300 python profiles/generic_md_workload.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:200]); continue
    print(d['case'][:28].ljust(28), d['ms_per_step'], d.get('specialized_ms_per_step'))"
