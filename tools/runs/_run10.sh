timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; python -c "
import json; r=json.load(open('gpurun_out/r02_bench_default.json')); print(r['value'], r['ms_per_step'], r['e2e']['ms_per_step'], r['roofline']['frac'], r['roofline']['traffic'], r['cpu_baseline'], r['parity'], r['gpu_launches'], r['clocks'])"
