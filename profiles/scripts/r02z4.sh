mkdir -p gpurun_out
for tp in 1024 512; do echo "== forced tile_px $tp"; EVREP_JIT_TILE_PX=$tp PYTHONPATH=$PWD timeout 300 python profiles/generic_md_workload.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:200]); continue
    print(d['case'][:28].ljust(28), d['ms_per_step'], d.get('specialized_ms_per_step'))"
done
