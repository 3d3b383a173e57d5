# Round-1 profile of the default bench (c3, 1 GPU, lower-block storage, append conditioning).
set -x
python bench.py --steps 32 --warmup 3 --no-cpu --no-both > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1b.csv \
    python bench.py --steps 32 --warmup 3 --no-cpu --no-both > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'rankb_|rank1_|pending_correct|cond_app_partial|cond_app_finish|cond_app_scalars|masked_argmax|step_vectors|itl_directed|gather_rows_sym' \
    --launch-skip 110 -c 40 -o gpurun_out/prof_r1b -f python bench.py --steps 32 --warmup 3 --no-cpu --no-both > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_r1b.ncu-rep --page raw --csv > gpurun_out/prof_r1b_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r1b.ncu-rep --page details > gpurun_out/prof_r1b_details.txt 2>/dev/null
ls -la gpurun_out/ | tail -8
