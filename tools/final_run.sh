# Round-end validation on one B200: GPU tests, the bench line (20-step and default), the ncu launch list and -- with a second
# argument "full" -- the full captures of the sweep kernels at BASELINE configs 2 / 3 / 4 (read afterwards with
# tools/record_traffic.py, tools/ncu_summary.py).  Usage (on the GPU box, from the repo root): bash tools/final_run.sh r03 [full]
T=${1:-r03}; O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; tail -2 $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --no-callers --no-extras --no-cpu-baseline > $O/${T}_bench_20steps.json 2> $O/${T}_bench_20steps.err
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; tail -c 300 $O/${T}_bench.json
Q="--no-cpu-baseline --no-e2e --no-callers --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 1 $Q > $O/${T}_ncu_launch.log 2>&1
[ "$2" = full ] || exit 0
F="--set full --clock-control none --import-source on -f"
ncu $F -k regex:pamr_sweep_duo -s 3 -c 1 -o $O/${T}_duo python bench.py --steps 2 --warmup 3 $Q > $O/${T}_ncu.log 2>&1
ncu $F -k regex:pamr_sweep_lattice -s 3 -c 1 -o $O/${T}_lat1 python bench.py --steps 2 --warmup 3 $Q >> $O/${T}_ncu.log 2>&1
ncu $F -k regex:pamr_sweep_duo -s 3 -c 1 -o $O/${T}_duo_coco python bench.py --workload coco_b16_c81_512 --steps 2 --warmup 3 $Q >> $O/${T}_ncu.log 2>&1
ncu $F -k regex:pamr_sweep_duo -s 3 -c 1 -o $O/${T}_duo_hires python bench.py --workload hires_b16_c21_1024 --steps 2 --warmup 3 $Q >> $O/${T}_ncu.log 2>&1
grep -c "Report" $O/${T}_ncu.log
