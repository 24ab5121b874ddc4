#!/bin/bash
# The bench lines committed under profiles/r02/ (one GPU; every line carries parity, roofline, cpu_baseline, e2e).
mkdir -p gpurun_out/bench
b() { TAG=$1; shift; python bench.py "$@" > gpurun_out/bench/$TAG.json 2> gpurun_out/bench/$TAG.err; cut -c1-230 gpurun_out/bench/$TAG.json; tail -1 gpurun_out/bench/$TAG.err | cut -c1-200; }
b bench_poisson_n1
b bench_poisson_p1_n1 --degree 1
b bench_poisson_p3_n1 --degree 3 --n 96
b bench_poisson_p4_n1 --degree 4 --n 48
b bench_elasticity96_n1 --workload elasticity
b bench_nurbs_p4_n1 --workload nurbs_p4
b bench_fcm64_n1 --workload fcm
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench/bench_reference_arm.json 2> gpurun_out/bench/bench_reference_arm.err; cut -c1-300 gpurun_out/bench/bench_reference_arm.json
