set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_packed.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_e.log 2>&1; echo pytest rc=$?
tail -3 gpurun_out/pytest_e.log
timeout 200 python bench.py --no-cpu --no-extras > gpurun_out/bench_e.json 2> gpurun_out/bench.err; echo bench rc=$?
cat gpurun_out/bench_e.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(round(d['value'],2),'Gev/s', d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', round(d['e2e']['value'],2), d['e2e']['soa9']['value'])"
tail -3 gpurun_out/bench.err
