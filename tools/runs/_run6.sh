python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 400 gpurun_out/r02_bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/r02_bench_reference.json
python bench.py --config c3 2>/dev/null | grep '^{' > gpurun_out/r02_bench_c3.json; python -c "
import json; r=json.load(open('gpurun_out/r02_bench_c3.json')); print('c3', r['value'], r['ms_per_step'], r['e2e']['ms_per_step'], r['parity'])"
python bench.py --config c4 2>/dev/null | grep '^{' > gpurun_out/r02_bench_c4.json; python -c "
import json; r=json.load(open('gpurun_out/r02_bench_c4.json')); print('c4', r['value'], r['ms_per_step'], r['e2e']['ms_per_step'], r['parity'])"
python bench.py --config c5 --steps 3 --warmup 2 2>/dev/null | grep '^{' > gpurun_out/r02_bench_c5_n1.json; python -c "
import json; r=json.load(open('gpurun_out/r02_bench_c5_n1.json')); print('c5n1', r['value'], r['ms_per_step'], r['e2e']['ms_per_step'] if r['e2e'] else None, r['parity'])"
python tools/secondary_bench.py > gpurun_out/r02_secondary_bench.json 2> gpurun_out/r02_secondary_bench.err; python -c "
import json
r=json.load(open('gpurun_out/r02_secondary_bench.json'))
for e in r['entries']: print(e['path'], 'wall',e['wall_ms'],'kernel',e['kernel_ms'],'floor',e['pcie_floor_ms'],e['parity'])"
ncu --set full --clock-control none --import-source on -k regex:gc_coef_refine -o gpurun_out/r02_prof_refine python tools/profile_workloads.py c2 > gpurun_out/r02_prof_refine.log 2>&1; tail -2 gpurun_out/r02_prof_refine.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_launches_bench.log 2>&1; tail -1 gpurun_out/r02_launches.csv | cut -c1-200
