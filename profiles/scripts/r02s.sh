mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -2
python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/bench_h.json').read())
print(round(d['value'],2), d['ms_per_step'], 'e2e', round(d['e2e']['value'],2), d['roofline']['kernel_ms'])
c=d['configs']; print('c2', c['config2_ergo12_gen1']['ms_per_step'], c['config2_ergo12_gen1']['eager_ms_per_step'], 'c3', c['config3_fused']['ms_per_step'])
print(d['parity_spot_check']['pass'])"
