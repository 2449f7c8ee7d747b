mkdir -p gpurun_out
cat > /tmp/c3.py <<'PY'
import torch, numpy as np
import event_representation_study_b200.batched as eb
from event_representation_study_b200.synth import device_batch
dev = torch.device("cuda", 0)
d = device_batch(32, 500_000, 720, 1280, dev, seed=5)
ev = eb.EventBatch(d["x"], d["y"], d["t"], d["p"], d["offsets"].cpu().numpy())
for _ in range(3):
    eb.order_ops_fused(ev, 720, 1280)
torch.cuda.synchronize()
PY
grep -n "def order_ops_fused" event_representation_study_b200/batched.py
for k in k_tore_tile_k k_time_surface_tile_s; do
ncu --set full --import-source on --clock-control none -k regex:$k -s 2 -c 1 -o gpurun_out/prof2_$k -f env PYTHONPATH=$PWD python /tmp/c3.py > gpurun_out/ncu2_$k.log 2>&1; tail -2 gpurun_out/ncu2_$k.log
done
