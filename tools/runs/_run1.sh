for k in 1 8; do
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gc_|adx_" -c 200 --csv --log-file gpurun_out/c5_launches_k$k.csv python bench.py --config c5 --files 4096 --steps 1 --warmup 0 --no-cpu --no-e2e --c5-chunks $k > /dev/null 2>&1
python - <<PY
import csv,io,collections
lines=[l for l in open('gpurun_out/c5_launches_k$k.csv') if l.startswith('"')]
tot=collections.OrderedDict()
for r in csv.DictReader(io.StringIO(''.join(lines))):
    if r['Metric Name']!='gpu__time_duration.sum': continue
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
    ms=v/1e6 if u in('ns','nsecond') else v/1e3 if u in ('us','usecond') else v
    name=r['Kernel Name'].split('(')[0][-40:]
    tot.setdefault(name,[]).append(round(ms,3))
print('chunks=$k')
for k_,v in tot.items(): print(k_, len(v), round(sum(v),2), v[:10])
PY
done
