for kb in 100 50 30; do
echo "--- EVREP_TILE_SMEM_KB=$kb"
EVREP_TILE_SMEM_KB=$kb python bench_extra.py --only config3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['workload'][:70], round(d['ms_per_step'],4), round(d['roofline']['frac'],3))
"
done
