#!/bin/bash
# Final ncu captures of the bench workloads (one GPU).  Reports land in gpurun_out/; read here with scripts/ncu_extract.py.
NCU="ncu --set full --clock-control none --import-source on"
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
$NCU -k regex:k_rows3d --launch-skip 2 -c 1 -o gpurun_out/r02f_rows $B > gpurun_out/ncu_rows.log 2>&1
$NCU -k regex:k_geom3d --launch-skip 2 -c 1 -o gpurun_out/r02f_geom $B > gpurun_out/ncu_geom.log 2>&1
$NCU -k regex:k_rows3d --launch-skip 6 -c 2 -o gpurun_out/r02f_elast $B --workload elasticity > gpurun_out/ncu_elast.log 2>&1
$NCU -k regex:k_assemble_elemset --launch-skip 1 -c 1 -o gpurun_out/r02f_fcm $B --workload fcm > gpurun_out/ncu_fcm.log 2>&1
$NCU -k regex:k_assemble_elemset --launch-skip 1 -c 1 -o gpurun_out/r02f_nurbs $B --workload nurbs_p4 > gpurun_out/ncu_nurbs.log 2>&1
$NCU -k regex:k_rows3d --launch-skip 2 -c 1 -o gpurun_out/r02f_rows_p3 $B --degree 3 --n 96 > gpurun_out/ncu_rows_p3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_poisson.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launches_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_elasticity.csv python bench.py --steps 2 --warmup 1 --no-cpu --workload elasticity > gpurun_out/launches_bench_e.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
tail -2 gpurun_out/ncu_*.log | cut -c1-200
