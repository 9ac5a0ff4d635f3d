python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/resample_probe.py 5e7 uniform_idx uniform lognormal > gpurun_out/probe.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum --clock-control none -k regex:"resample_rec" --csv --log-file gpurun_out/probe_ncu.csv python tools/resample_probe.py 5e7 uniform_idx uniform lognormal > gpurun_out/ncu.log 2>&1; cat gpurun_out/probe.log
python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'])"
