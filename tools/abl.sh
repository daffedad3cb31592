# A/B timing of library variants built with cl4wsis_b200.build.build(extra=..., lib=libcl4_<name>.so)
run() { echo -n "$1: "; CL4_LIB=$2 timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['roofline']['mean_launch_ms'],4), round(d['value'],1), d['checksums'])"; }
run base ""
for v in "$@"; do run $v $PWD/cl4wsis_b200/libcl4_$v.so; done
