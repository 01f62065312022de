python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; cat gpurun_out/bench_r2b.json | python -c "
import sys,json
r=json.loads(sys.stdin.read()); 
print('value',r['value'],'ms',r['ms_per_step'],'e2e',r['e2e']['value'],r['e2e']['ms_per_step'],'kernel',r['kernel_ms'],'tp',r['time_parallel'],'parity',r['parity'], 'roof', r['roofline']['frac'])
print(r['e2e']['timeline_ms'])"
tail -2 gpurun_out/bench_r2b.err
