set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo full bench rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/bench_full.json').readline())
print(round(d['value'],2),'Gev/s', d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', round(d['e2e']['value'],2), d['e2e']['soa9']['value'])
print('configs', d['configs']['config2_ergo12_gen1'])
print('dropin', {k:(v['ms_per_window'], v['speedup_vs_port']) for k,v in d['dropin'].items()})
"
tail -5 gpurun_out/bench_full.err
