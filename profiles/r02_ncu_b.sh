set -x
B="python bench.py --steps 1 --warmup 1 --configs none --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_short.csv $B > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"lookup_kernel|vote_bits_kernel|sketch_kernel" --launch-skip 36 --launch-count 3 -f -o gpurun_out/r02_ncu_short2 $B > gpurun_out/ncu_short2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lookup_kernel" --launch-skip 12 --launch-count 1 -f -o gpurun_out/r02_ncu_lookup_cold $B > gpurun_out/ncu_short3.log 2>&1
ls -la gpurun_out/*.ncu-rep
