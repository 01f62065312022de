python -m pytest tests/test_gcadpcm_segments_gpu.py tests/test_containers_gpu.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print('c2', r['ms_per_step'], r['kernel_ms'], 'e2e', r['e2e']['ms_per_step'], [g[1] for g in r['e2e']['timeline_ms']['groups']], r['e2e']['matches_device_resident'])"
for f in dsp; do python bench.py --config batch --out-format $f --steps 3 --warmup 2 2>/dev/null | grep '^{' > gpurun_out/r02_bench_batch_$f.json; python -c "
import json; r=json.load(open('gpurun_out/r02_bench_batch_$f.json')); print('$f', r['value'], r['ms_per_step'], r['stage_ms'], r['parity'])"; done
