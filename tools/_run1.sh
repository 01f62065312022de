set -x
python bench.py --config c3 --channels 256 --seconds 3 --steps 2 --warmup 1 2>&1 | tail -3 | cut -c1-1500
python bench.py --config c4 --channels 32 --seconds 3 --steps 2 --warmup 1 2>&1 | tail -3 | cut -c1-1500
python bench.py --config c5 --files 512 --steps 2 --warmup 1 2>&1 | tail -3 | cut -c1-1800
