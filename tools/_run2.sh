python -m pytest tests/test_multidevice_gpu.py tests/test_gcadpcm_gpu.py -m gpu -q -x 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config c5 --steps 3 --warmup 2 > gpurun_out/r02_bench_c5_n2.json 2> gpurun_out/r02_bench_c5_n2.err; tail -c 2200 gpurun_out/r02_bench_c5_n2.json; tail -3 gpurun_out/r02_bench_c5_n2.err
python bench.py --config c5 --steps 3 --warmup 2 --no-cpu > gpurun_out/r02_bench_c5_n1b.json 2> gpurun_out/r02_bench_c5_n1b.err; python -c "
import json; r=json.load(open('gpurun_out/r02_bench_c5_n1b.json')); print('n1', r['value'], r['ms_per_step'], r['e2e']['ms_per_step'])"
