python -m pytest tests/test_criadx_gpu.py -m gpu -q -x 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config c5 --steps 3 --warmup 2 > gpurun_out/r02_bench_c5_n2.json 2> gpurun_out/r02_bench_c5_n2.err; python -c "
import json; r=json.load(open('gpurun_out/r02_bench_c5_n2.json')); print('n2', r['value'], r['ms_per_step'], r['collective'], r['parity'], 'e2e', r['e2e']['ms_per_step'])"; tail -2 gpurun_out/r02_bench_c5_n2.err
python bench.py --config c5 --steps 3 --warmup 2 > gpurun_out/r02_bench_c5_n1.json 2> gpurun_out/r02_bench_c5_n1.err; python -c "
import json; r=json.load(open('gpurun_out/r02_bench_c5_n1.json')); print('n1', r['value'], r['ms_per_step'], r['collective']['encode_ms_per_rank'], r['parity'], 'e2e', r['e2e']['ms_per_step'])"
python tools/secondary_bench.py > gpurun_out/r02_secondary_bench.json 2> gpurun_out/r02_secondary_bench.err; python -c "
import json
r=json.load(open('gpurun_out/r02_secondary_bench.json'))
for e in r['entries']: print(e['path'], 'wall',e['wall_ms'],'kernel',e['kernel_ms'],'floor',e['pcie_floor_ms'],'ratio',e['wall_over_floor'],e['parity'])"
