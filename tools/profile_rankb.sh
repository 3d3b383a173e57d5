# one --set full capture of the deferred-update pass (rankb_kernel, 16 updates) of the default bench
ncu --set full --clock-control none --import-source on -k regex:'rankb_kernel' -c 1 -o gpurun_out/prof_rankb -f \
    python bench.py --steps 16 --warmup 3 --no-cpu --no-both > gpurun_out/ncu_rankb.log 2>&1
ncu -i gpurun_out/prof_rankb.ncu-rep --page raw --csv > gpurun_out/prof_rankb_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_rankb.ncu-rep --page details > gpurun_out/prof_rankb_details.txt 2>/dev/null
ls -la gpurun_out/prof_rankb*
