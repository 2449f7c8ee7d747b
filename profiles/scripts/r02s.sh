mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_packed.py -q 2>&1 | tail -5
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_packed.py -q -k "delta" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_f3.json 2> gpurun_out/bench_f3.err; echo rc=$?; tail -3 gpurun_out/bench_f3.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_f3.json').read()); e=d['e2e']
print(round(d['value'],2), d['ms_per_step'], 'e2e', round(e['value'],2), 'soa', round(e['soa9']['value'],2), e['h2d_bytes_per_step'], e['link'], e['host_format'][:60])"
