#!/bin/bash
# round-end evidence: default bench line, ncu launch list of the same command, one full capture of a whole step
set -x
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"predict_loglik|tile_|scan_|guide_|bucket_|record_|resample_|systematic" -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 2 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none -k regex:"predict_loglik|tile_|scan_|guide_|bucket_|record_|resample_" -s 27 -c 9 -o /tmp/prof_final -f python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/ncu2.log 2>&1
ncu -i /tmp/prof_final.ncu-rep --page raw --csv > gpurun_out/final_raw.csv 2>/dev/null
ls -la gpurun_out/
