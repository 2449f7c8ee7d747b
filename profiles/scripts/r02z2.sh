mkdir -p gpurun_out
PYTHONPATH=$PWD timeout 300 python profiles/generic_md_workload.py 2>&1 | tee gpurun_out/generic_md.jsonl | cut -c1-110
# launch list of config 2 (ERGO-12 Gen1 32 x 200 k)
cat > /tmp/c2.py <<'PY'
import torch
import event_representation_study_b200.batched as eb
from event_representation_study_b200.synth import device_batch
dev = torch.device("cuda", 0)
d = device_batch(32, 200_000, 240, 304, dev, seed=5)
ev = eb.EventBatch(d["x"], d["y"], d["t"], d["p"], d["offsets"].cpu().numpy())
for _ in range(4):
    eb.ergo12(ev, 240, 304)
torch.cuda.synchronize()
PY
PYTHONPATH=$PWD timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c2.csv python /tmp/c2.py > /dev/null 2>&1; echo ncu c2 rc=$?
grep -E "k_init|k_hist|k_colscan|k_scan|k_bin|k_md" gpurun_out/launches_c2.csv | tail -8 | cut -d, -f5,12- | cut -c1-200
