# timing-only ablations of the ERGO-12 tile kernel (results are wrong by construction): what each phase costs
# abl1 = accumulators not re-zeroed, abl2 = finalise without channel arithmetic (2 of 17 accumulator reads), abl3 = no accumulate phase, abl4 = no TMA store
mkdir -p gpurun_out
L=$PWD/event_representation_study_b200/lib
for v in std abl1 abl2 abl3 abl4; do
  f=$L/libevrep_$v.so; [ $v = std ] && f=$L/libevrep.so
  EVREP_LIB=$f timeout 120 python bench.py --no-cpu --no-extras --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$v', d['ms_per_step'], d['roofline']['kernel_ms']['tile kernel'])"
done | tee gpurun_out/z10.log
