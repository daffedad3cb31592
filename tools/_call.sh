python -m pytest tests -m gpu -x -q -k "pamr or sweep or lattice" 2>&1 | tail -2
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-callers --no-extras"
ncu --set full --clock-control none --import-source on -k regex:pamr_sweep_duo --launch-skip 20 -c 1 -f -o gpurun_out/r02s_duo $B > gpurun_out/r02s_ncu.log 2>&1
