mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mdjit.py -m gpu -x -q 2>&1 | tail -5
timeout 500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read())
print(round(d['value'],2), d['ms_per_step'], 'e2e', round(d['e2e']['value'],2), d['roofline']['frac'], d['roofline']['whole_step']['frac'])
print({k: round(v['ms_per_step'],4) for k,v in d['configs'].items()})
st=d['configs']['search_tuple']; print({k:v for k,v in st.items() if k not in ('windows','functions','aggregations')})
print(d['parity_spot_check'])"
tail -3 gpurun_out/bench.err
