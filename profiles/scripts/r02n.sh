mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?
wc -c gpurun_out/bench_n2.json
grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_n2.err | tail -20
python -c "
import json; d=json.loads(open('gpurun_out/bench_n2.json').readline())
print(d['n_gpus'], round(d['value'],2),'Gev/s', d['ms_per_step'], 'e2e', round(d['e2e']['value'],2), d['e2e']['soa9']['value'])
print('gwd', {k:v for k,v in d['gwd'].items() if k!='ranking'})
print(list(d.keys()))
"
