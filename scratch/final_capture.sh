#!/bin/bash
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1z_launches.csv python bench.py --steps 3 --warmup 1 --no-extras --no-cpu > gpurun_out/z0.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ss_fwd_rows_kernel -s 1 -c 1 -o gpurun_out/r1z_fwd -f python bench.py --steps 2 --warmup 1 --no-extras --no-cpu > gpurun_out/z1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ss_bwd_rows_kernel -s 1 -c 1 -o gpurun_out/r1z_bwd -f python bench.py --steps 2 --warmup 1 --no-extras --no-cpu > gpurun_out/z2.log 2>&1
timeout 400 python bench.py 2>gpurun_out/z4.err | tail -1 > gpurun_out/r1z_bench.json
ls -la gpurun_out/r1z_bench.json gpurun_out/r1z_fwd.ncu-rep gpurun_out/r1z_bwd.ncu-rep gpurun_out/r1z_launches.csv
