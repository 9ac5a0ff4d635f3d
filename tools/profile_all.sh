#!/bin/bash
# Evidence for every kernel family (profiles/README.md): one bench line per workload (no profiler), then one
# `ncu --set full` capture of the same workload's kernels; raw pages land in gpurun_out/, tools/ncu_summary.py
# condenses them into profiles/.   usage: gpurun --timeout 1500 -- bash tools/profile_all.sh <tag> [workloads...]
tag=$1; shift
K='predict|loglik|msort|scan_|tile_|wrap_|record_|resample|moments|lse_|weights_|systematic|guide|bucket|sharded'
for wl in "$@"; do
  extra=""; name=$wl
  case $wl in
    t3sys) extra="--workload t3 --resampling systematic";;
    t3mn) extra="--workload t3 --resampling multinomial";;
    standalone) ;;
    *) extra="--workload $wl";;
  esac
  if [ "$wl" = standalone ]; then
    python tools/standalone_kernels.py > gpurun_out/${tag}_standalone.log 2>&1 || { echo "standalone failed"; continue; }
    ncu --set full --clock-control none -k regex:"$K|rng_" -s 12 -c 12 -o /tmp/prof_$wl -f python tools/standalone_kernels.py > gpurun_out/ncu_$wl.log 2>&1
  else
    python bench.py $extra --steps 10 --warmup 3 > gpurun_out/${tag}_bench_$name.json 2> gpurun_out/${tag}_bench_$name.err || { echo "bench $wl failed"; tail -3 gpurun_out/${tag}_bench_$name.err; continue; }
    ncu --set full --clock-control none -k regex:"$K" -c 40 -o /tmp/prof_$wl -f python bench.py $extra --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/ncu_$wl.log 2>&1
  fi
  ncu -i /tmp/prof_$wl.ncu-rep --page raw --csv > gpurun_out/${tag}_raw_$name.csv 2>/dev/null
  echo "$wl done: $(wc -l < gpurun_out/${tag}_raw_$name.csv) rows"
done
