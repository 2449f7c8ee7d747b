# A/B: L2 cache hints - output bulk stores evict-first (ef), plus record stores evict-last (efl)
mkdir -p gpurun_out
L=$PWD/event_representation_study_b200/lib
for rep in 1 2; do
for v in std ef efl; do
  f=$L/libevrep_$v.so; [ $v = std ] && f=$L/libevrep.so
  EVREP_LIB=$f timeout 120 python bench.py --no-cpu --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$v', round(d['value'],2), 'Gev/s', d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', round(d['e2e']['value'],2))"
done; done | tee gpurun_out/z8.log
f=$L/libevrep_efl.so
EVREP_LIB=$f timeout 150 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_md_tile_static|k_bin" -s 8 -c 4 python bench.py --steps 2 --warmup 1 --no-cpu --no-extras 2>/dev/null | grep -E "k_md_tile_static|k_bin<|dram__bytes|gpu__time" | cut -c1-150
