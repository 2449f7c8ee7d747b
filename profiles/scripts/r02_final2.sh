# final evidence of round 2 (second session) on the committed tree: GPU tests, smoke, bench lines, reference arm, launch lists, ncu captures
# of the hot kernels (headline: tile, k_bin, k_hist; a run-time specialised tuple's tile kernel), mixed-density tuples interpreted vs
# specialised, compute-sanitizer over every kernel family
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/smoke.log
timeout 500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench rc=$?
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo ref rc=$?
timeout 200 python bench.py --clustered --no-cpu --no-extras > gpurun_out/bench_clustered.json 2>> gpurun_out/bench.err; echo clustered rc=$?
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_bench.log 2>&1; echo ncu launches rc=$?
for k in k_md_tile_static:tile k_bin:bin k_hist:hist; do
  kn=${k%%:*}; tag=${k##*:}
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:$kn -s 4 -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_$tag.log 2>&1; echo ncu $tag rc=$?
done
cat > /tmp/jit1.py <<'PY'
import torch
import event_representation_study_b200.batched as eb
from event_representation_study_b200.synth import device_batch
dev = torch.device("cuda", 0)
d = device_batch(32, 1_000_000, 720, 1280, dev, seed=3)
ev = eb.EventBatch(d["x"], d["y"], d["t"], d["p"], d["offsets"].cpu().numpy())
wi = [4, 2, 4, 6, 5, 1, 0, 4, 4, 5, 1, 2]
fu = ["timestamp", "timestamp_neg", "count_pos", "timestamp", "timestamp_neg", "timestamp", "timestamp_neg", "polarity", "timestamp_pos", "count_pos", "timestamp_neg", "timestamp_pos"]
ag = ["max", "variance", "variance", "max", "max", "mean", "mean", "mean", "sum", "max", "variance", "max"]
out = torch.empty((32, 720, 1280, 12), device=dev)
for _ in range(3):
    eb.mixed_density(ev, 720, 1280, wi, fu, ag, "SBN", out=out, specialize=True)
torch.cuda.synchronize()
PY
timeout 150 env PYTHONPATH=$PWD ncu --set full --clock-control none --import-source on -k regex:k_md_tile_static -s 1 -c 1 -f -o gpurun_out/prof_jit python /tmp/jit1.py > gpurun_out/ncu_jit.log 2>&1; echo ncu jit rc=$?
PYTHONPATH=$PWD timeout 300 python profiles/generic_md_workload.py > gpurun_out/generic_md.jsonl 2>&1; echo generic rc=$?
python profiles/sanitizer_workload.py 2>&1 | tail -1
(echo "# compute-sanitizer on profiles/sanitizer_workload.py (every kernel family once, small sizes, incl. the run-time specialised mixed-density kernels), B200, round 2 (final tree)"; echo "## memcheck"; timeout 600 compute-sanitizer --tool memcheck python profiles/sanitizer_workload.py 2>&1 | grep -E "ERROR SUMMARY|Invalid|error:" | head -8; echo "## racecheck"; timeout 900 compute-sanitizer --tool racecheck python profiles/sanitizer_workload.py 2>&1 | grep -E "RACECHECK SUMMARY|hazard" | head -8; echo "## synccheck"; timeout 600 compute-sanitizer --tool synccheck python profiles/sanitizer_workload.py 2>&1 | grep -E "ERROR SUMMARY|Barrier error" | head -8) > gpurun_out/r02_sanitizer.txt
cat gpurun_out/r02_sanitizer.txt
timeout 400 python bench_extra.py > gpurun_out/bench_extra.log 2> gpurun_out/bench_extra.err; echo extra rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read())
print(round(d['value'],2), d['ms_per_step'], 'e2e', round(d['e2e']['value'],2), d['roofline']['frac'], d['roofline']['whole_step']['frac'])
print({k: round(v['ms_per_step'],4) for k,v in d['configs'].items()}, d['parity_spot_check']['pass'], {k: round(v['ms_per_window'],2) for k,v in d['dropin'].items()})
print(d['gwd']['value'], d['gwd']['paper_shaped']['ms_per_pair'], d['cpu_baseline']['value'])"
cat gpurun_out/bench_ref.json | cut -c1-300
