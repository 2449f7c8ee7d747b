# final evidence of round 1: bench lines of every workload on the final tree, the reference arm, one full ncu capture of the
# dominant kernel (tile kernel) and of k_bin inside the bench command
mkdir -p gpurun_out
timeout 200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench rc=$?
timeout 250 python bench.py --impl reference > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo ref rc=$?
timeout 100 ncu --set full --clock-control none --import-source on -k regex:k_md_tile_static -s 4 -c 1 -f -o gpurun_out/prof_tile python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_tile.log 2>&1; echo ncu tile rc=$?
timeout 100 ncu --set full --clock-control none --import-source on -k regex:k_bin -s 4 -c 1 -f -o gpurun_out/prof_bin python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bin.log 2>&1; echo ncu bin rc=$?
timeout 400 python bench_extra.py > gpurun_out/bench_extra.log 2> gpurun_out/bench_extra.err; echo extra rc=$?
cat gpurun_out/bench.json gpurun_out/bench_ref.json | cut -c1-400
