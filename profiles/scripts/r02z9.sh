# A/B: accumulator slabs cleared with st.bulk (UMEMSETS) instead of 16-byte stores
mkdir -p gpurun_out
L=$PWD/event_representation_study_b200/lib
for rep in 1 2; do
for v in std zb; do
  f=$L/libevrep_$v.so; [ $v = std ] && f=$L/libevrep.so
  EVREP_LIB=$f timeout 120 python bench.py --no-cpu --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$v', round(d['value'],2), 'Gev/s', d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', round(d['e2e']['value'],2))"
done; done | tee gpurun_out/z9.log
EVREP_LIB=$L/libevrep_zb.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ergo" 2>&1 | tail -2
