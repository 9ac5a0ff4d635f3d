#!/bin/bash
# Round-end evidence on one GPU (run under gpurun): the default bench line and the reference arm, the ncu launch list
# of the same command, one `ncu --set full` capture of a whole step, and one bench line per other workload.
# Summaries for profiles/: python tools/ncu_summary.py gpurun_out/<tag>_step_raw.csv profiles/<tag>_ncu_full_T3_1e8_summary.csv
#                          python tools/traffic_from_summary.py t3:100000000:f64:multinomial_sorted:exact=profiles/<tag>_ncu_full_T3_1e8_summary.csv
tag=${1:-final}
K='predict|loglik|msort|scan_|tile_|wrap_|record_|resample|moments|lse_|weights_|systematic|guide|bucket|sharded'
set -x
python bench.py > gpurun_out/${tag}_bench_t3.json 2> gpurun_out/${tag}_bench_t3.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 200 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 2 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"predict|loglik|msort|scan_|tile_" -s 27 -c 9 -o /tmp/prof_step -f python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/ncu_step.log 2>&1
ncu -i /tmp/prof_step.ncu-rep --page raw --csv > gpurun_out/${tag}_step_raw.csv 2>/dev/null
for w in s2 r4 t2mc r4s r2; do python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err; done
python bench.py --resampling systematic --steps 10 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_t3sys.json 2>/dev/null
python bench.py --resampling multinomial --steps 10 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_t3mn.json 2>/dev/null
ls -la gpurun_out/
