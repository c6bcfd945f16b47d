timeout 220 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/pytest_r2zz.log; cat gpurun_out/pytest_r2zz.log | tail -2
timeout 30 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r2zz.log 2>&1; tail -1 gpurun_out/smoke_r2zz.log
timeout 170 python bench.py > gpurun_out/bench_r2zz.json 2> gpurun_out/bench_r2zz.err; tail -c 300 gpurun_out/bench_r2zz.json
timeout 70 ncu --set full --clock-control none --import-source on -k 'regex:ss_(fwd|bwd)_cw_kernel' -c 4 -f -o gpurun_out/r2zz_scan python bench.py --steps 1 --warmup 1 --no-extras --no-cpu --no-modules > gpurun_out/ncu_r2zz.log 2>&1; tail -2 gpurun_out/ncu_r2zz.log
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2zz_launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu --no-modules > /dev/null 2>&1; wc -l gpurun_out/r2zz_launches.csv
