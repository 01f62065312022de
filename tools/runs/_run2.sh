python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config c5 --steps 3 --warmup 2 --no-cpu 2>gpurun_out/r02_bench_c5_n2.err | grep '^{' > gpurun_out/r02_bench_c5_n2.json; python -c "
import json; r=json.load(open('gpurun_out/r02_bench_c5_n2.json')); print('n2', r['value'], r['ms_per_step'], r['collective'], 'e2e', r['e2e']['ms_per_step'])"; tail -2 gpurun_out/r02_bench_c5_n2.err
python bench.py --config c5 --steps 3 --warmup 2 --no-cpu --no-e2e 2>gpurun_out/r02_bench_c5_n1.err | grep '^{' > gpurun_out/r02_bench_c5_n1.json; python -c "
import json; r=json.load(open('gpurun_out/r02_bench_c5_n1.json')); print('n1', r['value'], r['ms_per_step'], r['collective']['encode_ms_per_rank'])"
