python -m pytest tests/test_gcadpcm_segments_gpu.py tests/test_gcadpcm_gpu.py -q 2>&1 | tail -5
python tools/seg_sweep.py 1024 30 1,17,24,34,0 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(r['segments'], r['gc_encode_ms'], r['fallback_frac'], r['cascade_boundaries'], r['same_bytes_as_first'])"
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; cat gpurun_out/bench_r2a.json; tail -3 gpurun_out/bench_r2a.err
