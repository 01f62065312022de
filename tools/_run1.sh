python -m pytest tests/test_gcadpcm_segments_gpu.py tests/test_gcadpcm_gpu.py -q -x 2>&1 | tail -3
for b in 8 10 12; do
  (cd vgaudio_b200/csrc && rm -f gc_encode.o && make EXTRA=-DVGB_ENC_BLOCKS_PER_SM=$b >/dev/null 2>&1)
  echo "BLOCKS=$b"
  python tools/seg_sweep.py 1024 30 1,24,34,48 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(r['segments'], r['gc_encode_ms'], r['fallback_frac'], r['cascade_boundaries'], r['same_bytes_as_first'])"
done
(cd vgaudio_b200/csrc && rm -f gc_encode.o && make >/dev/null 2>&1)
