python tools/seg_sweep.py 1024 30 24,0 2>&1 | tail -2
