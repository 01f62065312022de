python -m pytest tests/test_containers_gpu.py tests/test_cli_gpu.py -m gpu -x -q 2>&1 | tail -2
for f in dsp adx; do python bench.py --config batch --out-format $f --steps 3 --warmup 2 --no-cpu 2>/dev/null | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print('$f', r['value'], r['ms_per_step'], r['stage_ms'], r['roofline']['byte_movers_gbs'])"; done
