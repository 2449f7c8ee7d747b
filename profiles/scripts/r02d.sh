set -x
mkdir -p gpurun_out
EVREP_SORTBIN_TEST=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ergo12" > gpurun_out/pytest_sortbin.log 2>&1; echo pytest rc=$?
tail -3 gpurun_out/pytest_sortbin.log
for v in 1 ""; do
EVREP_SORTBIN_TEST=$v timeout 200 python bench.py --no-cpu --no-extras > gpurun_out/bench_sb.json 2> gpurun_out/bench.err; echo bench rc=$?
cat gpurun_out/bench_sb.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('SORTBIN_TEST=$v', round(d['value'],2),'Gev/s', d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', round(d['e2e']['value'],2), d['e2e']['soa9']['value'])"
done
EVREP_SORTBIN_TEST=1 timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_sortbin -s 4 -c 1 -f -o gpurun_out/prof_k_sortbin python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_sortbin.log 2>&1; echo ncu rc=$?
