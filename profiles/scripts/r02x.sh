for g in 1 2 3 4 8; do
python bench.py --steps 20 --warmup 3 --no-cpu --no-extras --e2e-groups $g 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('groups', $g, 'e2e', round(e['value'],2), 'soa', round(e['soa9']['value'],2), 'frac', round(e['link']['frac_of_link'],3), 'link', round(e['link']['h2d_gbs_per_gpu_all_ranks_copying'],1))"
done
