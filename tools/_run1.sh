python -m pytest tests/test_criadx_gpu.py tests/test_multidevice_gpu.py tests/test_cri_golden.py -m gpu -q -x 2>&1 | tail -5
python tools/secondary_bench.py > gpurun_out/r02_secondary_bench.json 2> gpurun_out/r02_secondary_bench.err; python -c "
import json
r=json.load(open('gpurun_out/r02_secondary_bench.json'))
for e in r['entries']: print(e['path'], 'wall',e['wall_ms'],'kernel',e['kernel_ms'],'floor',e['pcie_floor_ms'],'ratio',e['wall_over_floor'],'roof',e['roofline']['frac'],'cpu',e['cpu_baseline']['value'],e['cpu_baseline']['cores'],e['parity'])"
tail -3 gpurun_out/r02_secondary_bench.err
