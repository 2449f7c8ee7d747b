mkdir -p gpurun_out
for g in 1 2 8 16; do
timeout 200 python bench.py --no-cpu --no-extras --e2e-groups $g > gpurun_out/bench_g.json 2> gpurun_out/bench.err; echo bench rc=$?
cat gpurun_out/bench_g.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('groups $g', 'e2e', round(d['e2e']['value'],2), round(d['e2e']['soa9']['value'],2))"
done
python - <<'PY'
import torch, time
dev=torch.device('cuda')
h=torch.empty(32*1024*1024, dtype=torch.int32).pin_memory()
d=[torch.empty_like(h, device=dev) for _ in range(4)]
st=torch.cuda.Stream()
for n in (1,4):
    torch.cuda.synchronize(); t0=time.perf_counter()
    with torch.cuda.stream(st):
        for r in range(5):
            for k in range(n):
                d[k].copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print('H2D', n, 'x128MB chunks:', 5*n*h.numel()*4/dt/1e9, 'GB/s')
PY
