set -x
B="python bench.py --steps 1 --warmup 1 --configs none --no-cpu-baseline"
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"lookup_kernel|vote_bits_kernel|sketch_kernel" --launch-skip 36 --launch-count 3 -f -o gpurun_out/r02_ncu_short4 $B > gpurun_out/ncu_short4.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"lookup_kernel|vote_bits_kernel|sketch_kernel" --launch-skip 30 --launch-count 5 -f -o gpurun_out/r02_ncu_multik4 $B --workload multik > gpurun_out/ncu_multik4.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"lookup_kernel|vote_long_kernel|sketch_kernel|vote_kernel" --launch-skip 64 --launch-count 4 -f -o gpurun_out/r02_ncu_long4 $B --workload long > gpurun_out/ncu_long4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sketch_kernel|lookup_kernel|vote_|radix_|scan_|compact_|class_|em_|seg_|as_partial|items_|fixed_layout|permute_out|make_sort|split_keys|fill_u32|all_to_slow" --csv --log-file gpurun_out/r02_launches_short4.csv $B > gpurun_out/ncu_l4.log 2>&1
