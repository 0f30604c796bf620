# round-2 final captures (run under gpurun, one GPU); every command exits 0 without ncu first (bench runs before this)
set -x
B="python bench.py --steps 1 --warmup 1 --configs none --no-cpu-baseline"
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"lookup_kernel|vote_bits_kernel|sketch_kernel" --launch-skip 36 --launch-count 3 -f -o gpurun_out/r02_final_short $B > gpurun_out/ncu_final_short.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"radix_scatter|radix_hist|class_head|class_gather|em_den_kernel|em_partial_kernel|em_mstep" --launch-skip 60 --launch-count 14 -f -o gpurun_out/r02_final_finish $B > gpurun_out/ncu_final_finish.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"lookup_kernel|vote_bits_kernel|sketch_kernel" --launch-skip 30 --launch-count 5 -f -o gpurun_out/r02_final_multik $B --workload multik > gpurun_out/ncu_final_multik.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"lookup_kernel|vote_long_kernel|sketch_kernel|vote_kernel" --launch-skip 64 --launch-count 4 -f -o gpurun_out/r02_final_long $B --workload long > gpurun_out/ncu_final_long.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sketch_kernel|lookup_kernel|vote_|radix_|scan_|read_keys|csr_|class_|em_|seg_|as_partial|items_|fixed_layout|permute_out|make_sort|split_keys|fill_u32|all_to_slow" --csv --log-file gpurun_out/r02_final_launches.csv $B > gpurun_out/ncu_final_l.log 2>&1
