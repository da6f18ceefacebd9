run() { echo "== $*"; env "$@" timeout 300 python bench.py --n-items 1250000 --steps 20 --warmup 5 --no-cpu-baseline --no-check 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['roofline']['appends_per_launch'], d['clocks']['sm_mhz'])"; }
run A=1
run BB_NO_PILOT=1
run BB_SPLITS=23
run BB_SPLITS=12
run BB_SPLITS=33
run BB_DEBUG_FLAGS=1
run BB_DEBUG_FLAGS=8
run BB_DEBUG_FLAGS=4
