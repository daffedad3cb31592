run() { echo -n "$1 dbg=$2: "; CL4_LIB=$3 CL4_DBG=$2 timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['roofline']['mean_launch_ms'],4))"; }
for d in 0 4 5 7 37 36; do run s4 $d ""; done
