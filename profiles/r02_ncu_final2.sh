# refresh of the config-2 captures after the last sketch-kernel changes (run under gpurun, one GPU)
set -x
B="python bench.py --steps 1 --warmup 1 --configs none --no-cpu-baseline"
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"lookup_kernel|vote_bits_kernel|sketch_kernel" --launch-skip 36 --launch-count 3 -f -o gpurun_out/r02_final_short $B > gpurun_out/ncu_final_short.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"radix_scatter|radix_hist|class_head|class_gather|class_fill" --launch-skip 16 --launch-count 12 -f -o gpurun_out/r02_final_sort $B > gpurun_out/ncu_final_sort.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sketch_kernel|lookup_kernel|vote_|radix_|scan_|read_keys|csr_|class_|em_|seg_|as_partial|items_|fixed_layout|permute_out|make_sort|split_keys|fill_u32|all_to_slow" --csv --log-file gpurun_out/r02_final_launches.csv $B > gpurun_out/ncu_final_l.log 2>&1
