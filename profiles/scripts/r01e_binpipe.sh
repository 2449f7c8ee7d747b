# A/B of the k_bin variants: 512 threads (sub-chunk fetched when needed) vs 256 threads with register double buffering
mkdir -p gpurun_out
for nt in 512 256 512 256; do
  EVREP_BIN_NT=$nt timeout 120 python bench.py --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('NT=$nt', round(d['value'],2), 'Gev/s', d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', round(d['e2e']['value'],2))"
done | tee gpurun_out/binpipe.log
EVREP_BIN_NT=256 timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_filters.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -3
