mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_nimagenet.py -x -q 2>&1 | tail -5
python - <<'PY'
import json, bench, torch
dev = torch.device("cuda", 0)
r = bench.bench_configs(dev, 20)
for k, v in r.items():
    print(k, round(v["ms_per_step"], 4), round(v["roofline"]["frac"], 3), v.get("eager_ms_per_step"), json.dumps(v.get("parity_spot_check"))[:300])
PY
python bench_extra.py --only config3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['workload'][:70], round(d['ms_per_step'],4), round(d['roofline']['frac'],3))
"
