timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -k "gwd or otmi or OTMI" 2>&1 | tail -3
python - <<'PY'
import json, bench, torch
dev = torch.device("cuda", 0)
r = bench.bench_gwd(0, 1, dev, 5, with_cpu=True)
print({k: v for k, v in r.items() if k not in ("ranking", "config")})
PY
