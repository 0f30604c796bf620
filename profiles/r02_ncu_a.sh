set -x
B="python bench.py --steps 1 --warmup 1 --configs none --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:"vote_bits_kernel|sketch_kernel" --launch-skip 24 --launch-count 2 -f -o gpurun_out/r02_ncu_short $B > gpurun_out/ncu_short.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"vote_bits_kernel" --launch-skip 12 --launch-count 1 -f -o gpurun_out/r02_ncu_multik $B --workload multik > gpurun_out/ncu_multik.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"vote_long_kernel" --launch-skip 20 --launch-count 1 -f -o gpurun_out/r02_ncu_long $B --workload long > gpurun_out/ncu_long.log 2>&1
ls -la gpurun_out/*.ncu-rep
