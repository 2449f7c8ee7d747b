cat > /tmp/c2.py <<'PY'
import torch, time
import event_representation_study_b200.batched as eb
from event_representation_study_b200 import _lib
from event_representation_study_b200.synth import device_batch
dev = torch.device("cuda", 0)
H, W, B, N = 240, 304, 32, 200_000
d = device_batch(B, N, H, W, dev, seed=2)
ev = eb.EventBatch(d["x"], d["y"], d["t"], d["p"], d["offsets"].cpu().numpy())
out = torch.empty((B, H, W, 12), device=dev)
for _ in range(5): eb.ergo12(ev, H, W, out=out)
torch.cuda.synchronize()
_lib.profile_enable(50)
for _ in range(50): eb.ergo12(ev, H, W, out=out)
torch.cuda.synchronize()
for k in (_lib.K_COUNT, _lib.K_SCAN, _lib.K_BIN, _lib.K_TILE):
    ms, n = _lib.profile_read(k); print(_lib.KERNEL_NAMES[k], round(ms / max(n, 1) * 1e3, 2), "us")
_lib.profile_enable(0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): eb.ergo12(ev, H, W, out=out)
e1.record(); torch.cuda.synchronize()
print("eager us/step", e0.elapsed_time(e1) / 200 * 1e3)
PY
PYTHONPATH=$PWD python /tmp/c2.py
PYTHONPATH=$PWD ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv python /tmp/c2.py 2>/dev/null | python -c "
import csv,sys,collections
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
h=rows[0]; agg=collections.OrderedDict()
for r in rows[1:]:
    d=dict(zip(h,r))
    if d.get('Metric Name')=='gpu__time_duration.sum' and 'evrep' in d['Kernel Name']:
        agg.setdefault(d['Kernel Name'][:50]+' '+d['Grid Size'],[]).append(float(d['Metric Value'].replace(',',''))/1e3)
for k,v in agg.items(): print(k, len(v), round(sum(v)/len(v),2),'us')
"
