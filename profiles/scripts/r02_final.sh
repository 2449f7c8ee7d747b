# final evidence of round 2 on the committed tree: GPU tests, smoke, bench lines of every workload, the reference arm, the ncu launch
# list of the bench command, and one full ncu capture of each hot kernel (headline: tile, k_bin, k_hist; config 3: the three order ops)
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench rc=$?
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo ref rc=$?
timeout 200 python bench.py --clustered --no-cpu --no-extras > gpurun_out/bench_clustered.json 2>> gpurun_out/bench.err; echo clustered rc=$?
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_bench.log 2>&1; echo ncu launches rc=$?
for k in k_md_tile_static:tile k_bin:bin k_hist:hist; do
  kn=${k%%:*}; tag=${k##*:}
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:$kn -s 4 -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_$tag.log 2>&1; echo ncu $tag rc=$?
done
cat > /tmp/c3.py <<'PY'
import torch
import event_representation_study_b200.batched as eb
from event_representation_study_b200.synth import device_batch
dev = torch.device("cuda", 0)
d = device_batch(32, 500_000, 720, 1280, dev, seed=5)
ev = eb.EventBatch(d["x"], d["y"], d["t"], d["p"], d["offsets"].cpu().numpy())
for _ in range(3):
    eb.order_ops_fused(ev, 720, 1280)
torch.cuda.synchronize()
PY
for k in k_tore_tile_k:tore k_time_surface_tile_s:ts k_event_stack_tile_k:es; do
  kn=${k%%:*}; tag=${k##*:}
  timeout 150 env PYTHONPATH=$PWD ncu --set full --clock-control none --import-source on -k regex:$kn -s 2 -c 1 -f -o gpurun_out/prof_$tag python /tmp/c3.py > gpurun_out/ncu_$tag.log 2>&1; echo ncu $tag rc=$?
done
timeout 200 env PYTHONPATH=$PWD ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c3.csv python /tmp/c3.py > /dev/null 2>&1; echo ncu c3 launches rc=$?
timeout 600 python bench_extra.py > gpurun_out/bench_extra.log 2> gpurun_out/bench_extra.err; echo extra rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read())
print(round(d['value'],2), d['ms_per_step'], 'e2e', round(d['e2e']['value'],2), d['roofline']['frac'], d['roofline']['whole_step']['frac'])
print({k: round(v['ms_per_step'],4) for k,v in d['configs'].items()}, d['parity_spot_check']['pass'], {k: round(v['ms_per_window'],2) for k,v in d['dropin'].items()})
print(d['gwd']['value'], d['gwd']['paper_shaped']['ms_per_pair'], d['cpu_baseline']['value'])"
cat gpurun_out/bench_ref.json | cut -c1-300
