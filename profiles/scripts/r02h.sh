mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c3.csv python bench_extra.py --only config3 --steps 3 > gpurun_out/c3.log 2>&1; echo rc=$?
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_c3.csv')) if len(r)>5]
hdr=rows[0]; iK=hdr.index('Kernel Name'); iV=hdr.index('Metric Value'); iG=hdr.index('Grid Size')
agg=collections.OrderedDict()
for r in rows[1:]:
    k=(r[iK][:70], r[iG]); v=float(r[iV].replace(',',''))
    agg.setdefault(k,[]).append(v)
for k,v in agg.items(): print(k, len(v), round(sum(v)/len(v)/1000,1),'us')
PY
for k in k_time_surface_tile_s k_tore_tile_k k_event_stack_tile_k; do
timeout 150 ncu --set full --clock-control none --import-source on -k regex:$k -s 13 -c 1 -f -o gpurun_out/prof_$k python bench_extra.py --only config3 --steps 3 > gpurun_out/ncu_$k.log 2>&1; echo ncu $k rc=$?
done
