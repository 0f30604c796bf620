set -x
B="python bench.py --steps 1 --warmup 1 --configs none --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sketch_kernel|lookup_kernel|vote_|radix_|scan_|read_keys|csr_|class_|em_|seg_|as_partial|items_|fixed_layout|permute_out|make_sort|split_keys|fill_u32|all_to_slow" --csv --log-file gpurun_out/r02_launches_short5.csv $B > gpurun_out/ncu_l5.log 2>&1
