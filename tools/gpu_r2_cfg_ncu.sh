#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --metrics lts__t_sector_hit_rate.pct,dram__sectors_read.sum,dram__sectors_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none --csv --log-file gpurun_out/r02_configs_ncu.csv python tools/run_configs.py 4 0 > gpurun_out/r02_configs_under_ncu.json 2> gpurun_out/r02_configs_under_ncu.err
tail -3 gpurun_out/r02_configs_under_ncu.err | cut -c1-300; wc -l gpurun_out/r02_configs_ncu.csv
