python -m pytest tests/test_crihca_gpu.py tests/test_containers_gpu.py -m gpu -x -q 2>&1 | tail -2
python bench.py --config c4 2>/dev/null | grep '^{' > gpurun_out/r02_bench_c4.json; python -c "
import json; r=json.load(open('gpurun_out/r02_bench_c4.json')); print('c4', r['value'], r['ms_per_step'], r['e2e']['ms_per_step'], r['parity'])"
