python bench.py --config c5 --steps 3 --warmup 2 --no-cpu > gpurun_out/r02_bench_c5_n1c.json 2> gpurun_out/r02_bench_c5_n1c.err; python -c "
import json; r=json.load(open('gpurun_out/r02_bench_c5_n1c.json')); print('c5 n1', r['value'], r['ms_per_step'], 'e2e', r['e2e']['ms_per_step'], r['collective'])"
tail -2 gpurun_out/r02_bench_c5_n1c.err
