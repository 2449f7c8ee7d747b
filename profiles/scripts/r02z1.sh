# A/B of k_bin / k_hist builds on one box: base = HEAD before the change; new = two register slots + t_before by shuffle;
# sc3 = super-chunks of 12288 events; c3 = 3 CTAs/SM (one slot, 16-bit cursors); sc3p = sc3 + 16-bit cursors
mkdir -p gpurun_out
L=$PWD/event_representation_study_b200/lib
for rep in 1 2; do
for v in base new sc3 c3 sc3p; do
  f=$L/libevrep_$v.so; [ $v = new ] && f=$L/libevrep.so
  EVREP_LIB=$f timeout 120 python bench.py --no-cpu --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$v', round(d['value'],2), 'Gev/s', d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', round(d['e2e']['value'],2))"
done; done | tee gpurun_out/z1.log
for v in new sc3 c3; do
  f=$L/libevrep_$v.so; [ $v = new ] && f=$L/libevrep.so
  EVREP_LIB=$f timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_filters.py -m gpu -x -q 2>&1 | tail -2
done
