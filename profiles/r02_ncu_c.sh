set -x
B="python bench.py --steps 1 --warmup 1 --configs none --no-cpu-baseline"
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"lookup_kernel|vote_bits_kernel|sketch_kernel" --launch-skip 36 --launch-count 3 -f -o gpurun_out/r02_ncu_short3 $B > gpurun_out/ncu_short3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sq::|sketch_kernel|lookup_kernel|vote_|radix_|scan_|compact_|class_|em_|seg_|as_partial|items_|fixed_layout|permute_out|make_sort|split_keys|fill_u32" --csv --log-file gpurun_out/r02_launches_short.csv $B > gpurun_out/ncu_l.log 2>&1
