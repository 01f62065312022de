python -m pytest tests/test_gcadpcm_gpu.py tests/test_gcadpcm_segments_gpu.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print('default', r['ms_per_step'], r['kernel_ms'], 'e2e', r['e2e']['ms_per_step'], r['e2e']['timeline_ms']['coefs_done'], [g[1] for g in r['e2e']['timeline_ms']['groups']], r['e2e']['matches_device_resident'])"
VGB_REFINE_WIDE_LIMIT=4096 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print('wide@1024', r['ms_per_step'], r['kernel_ms'])"
VGB_REFINE_WIDE_LIMIT=0 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print('narrow', r['ms_per_step'], 'e2e', r['e2e']['ms_per_step'])"
