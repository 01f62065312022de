N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --config c5 --steps 3 --warmup 2 $2 2>gpurun_out/r02_bench_c5_n${N}$3.err | grep '^{' > gpurun_out/r02_bench_c5_n${N}$3.json; python -c "
import json; r=json.load(open('gpurun_out/r02_bench_c5_n${N}$3.json')); print('n$N', r['value'], r['ms_per_step'], r['collective'], r['parity'], r['e2e'])"; tail -2 gpurun_out/r02_bench_c5_n${N}$3.err | cut -c1-300
