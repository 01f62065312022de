timeout 1500 python -m pytest tests/test_containers_gpu.py tests/test_cli_gpu.py -m gpu -x -q 2>&1 | tail -5
for f in dsp adx hca; do python bench.py --config batch --out-format $f --steps 3 --warmup 2 2>gpurun_out/r02_bench_batch_$f.err | grep '^{' > gpurun_out/r02_bench_batch_$f.json; python -c "
import json; r=json.load(open('gpurun_out/r02_bench_batch_$f.json')); print('$f', r['value'], r['ms_per_step'], r['stage_ms'], r['roofline']['byte_movers_gbs'], r['cpu_baseline']['value'] if r['cpu_baseline'] else None, r['parity'])"; tail -2 gpurun_out/r02_bench_batch_$f.err; done
