# A/B timing of library variants built with tools/build_variants.py (libcl4_<name>.so); CL4_SWEEP is passed through
run() { echo -n "$1: "; CL4_LIB=$2 timeout 120 python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --no-e2e --no-callers --no-extras --workload ${WORKLOAD:-voc_b16_c21_512} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['roofline']['mean_launch_ms'],4), round(d['value'],1), d['checksums'])"; }
run base ""
for v in "$@"; do run $v $PWD/cl4wsis_b200/libcl4_$v.so; done
