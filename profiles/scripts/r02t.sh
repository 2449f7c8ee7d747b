mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m 2>/dev/null | head -14
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo rc=$?
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; echo rc=$?
  fi
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/scale_n*.json')):
    try:
        d = json.loads(open(f).read())
    except Exception as e:
        print(f, 'unparsable', e); continue
    e = d['e2e']
    print(d['n_gpus'], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'e2e', round(e['value'], 2), 'soa9', round(e['soa9']['value'], 2), 'link', {k: (round(v, 2) if isinstance(v, float) else v) for k, v in e['link'].items() if k != 'note'}, e['placement'], 'gwd', (d.get('gwd') or {}).get('value'))
PY
