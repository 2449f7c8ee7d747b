# round 2, first GPU call: new tile-kernel epilogue (warp-local repack + TMA store, 2 barriers per bucket) and the lean k_hist / k_bin
set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/pytest_gpu.log
timeout 200 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench rc=$?
cat gpurun_out/bench.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(round(d['value'],2),'Gev/s', d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', round(d['e2e']['value'],2), 'frac', d['roofline']['whole_step']['frac'])"
timeout 200 python bench.py --no-cpu --clustered > gpurun_out/bench_clustered.json 2>> gpurun_out/bench.err; echo bench rc=$?
timeout 150 compute-sanitizer --tool racecheck python profiles/sanitizer_workload.py > gpurun_out/racecheck.log 2>&1; echo racecheck rc=$?; tail -4 gpurun_out/racecheck.log
for k in k_md_tile_static k_bin k_hist; do
timeout 150 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_$k.log 2>&1; echo ncu $k rc=$?
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo ncu rc=$?
timeout 300 python bench_extra.py > gpurun_out/bench_extra.log 2> gpurun_out/bench_extra.err; echo extra rc=$?
tail -30 gpurun_out/bench_extra.log
