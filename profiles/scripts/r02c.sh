set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/pytest_gpu.log
timeout 200 python bench.py --no-cpu --no-extras > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench rc=$?
cat gpurun_out/bench.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(round(d['value'],2),'Gev/s', d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', round(d['e2e']['value'],2), 'frac', d['roofline']['whole_step']['frac'])"
for k in $NCU_KERNELS; do
timeout 150 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_$k.log 2>&1; echo ncu $k rc=$?
done
if [ -n "$FULL_BENCH" ]; then
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo full bench rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/bench_full.json').readline())
print({k:(v if not isinstance(v,(dict,list)) else '...') for k,v in d.items()})
print('gwd', {k:v for k,v in d['gwd'].items() if k!='ranking'})
print('parity', d['parity_spot_check'])
print('configs', d['configs'])
print('dropin', d['dropin'])
print('cpu', d.get('cpu_baseline'))
"
tail -5 gpurun_out/bench_full.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?; cut -c1-300 gpurun_out/bench_ref.json
fi
