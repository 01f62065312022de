timeout 1500 python -m pytest tests/test_containers_gpu.py tests/test_cli_gpu.py -m gpu -x -q 2>&1 | tail -3
for f in dsp adx hca; do python bench.py --config batch --out-format $f --steps 3 --warmup 2 2>gpurun_out/r02_bench_batch_$f.err | grep '^{' > gpurun_out/r02_bench_batch_$f.json; python -c "
import json; r=json.load(open('gpurun_out/r02_bench_batch_$f.json')); print('$f', r['value'], r['ms_per_step'], r['stage_ms'], r['roofline']['byte_movers_gbs'], r['cpu_baseline']['value'] if r['cpu_baseline'] else None, r['parity'])"; tail -2 gpurun_out/r02_bench_batch_$f.err; done
python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print('narrow', r['ms_per_step'], r['kernel_ms'], 'e2e', r['e2e']['ms_per_step'], r['e2e']['timeline_ms']['coefs_done'])"
VGB_REFINE_WIDE_LIMIT=444 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print('wide<=444', r['ms_per_step'], r['kernel_ms'], 'e2e', r['e2e']['ms_per_step'], r['e2e']['timeline_ms']['coefs_done'])"
VGB_REFINE_WIDE_LIMIT=444 python bench.py --config batch --out-format dsp --steps 3 --warmup 2 --no-cpu 2>/dev/null | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print('batch dsp wide', r['ms_per_step'])"
