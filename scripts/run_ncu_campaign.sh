#!/bin/bash
# Final ncu captures of the bench workloads (one GPU).  The reports are reduced ON THE BOX to their raw-metric CSV, the details page and
# the per-stage stall table (the reports themselves exceed what a call may bring back); scripts/ncu_extract.py reads the CSVs.
NCU="ncu --set full --clock-control none --import-source on"
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
O=gpurun_out/ncu
mkdir -p $O
cap() { # tag kernel-regex launch-skip count bench-args...
  TAG=$1; K=$2; SKIP=$3; CNT=$4; shift 4
  $NCU -k regex:$K --launch-skip $SKIP -c $CNT -o /tmp/$TAG $B "$@" > $O/$TAG.log 2>&1
  ncu -i /tmp/$TAG.ncu-rep --page raw --csv > $O/${TAG}_raw.csv 2>/dev/null
  ncu -i /tmp/$TAG.ncu-rep --page details > $O/${TAG}_details.txt 2>/dev/null
}
cap rows k_rows3d 2 1
python scripts/ncu_stages.py /tmp/rows.ncu-rep nutils_b200/csrc/assemble_rows.cu "---- S1: contract Q2" "---- S2: contract Q1" "// Out-of-line variants" "// The kernel." "---- work unit ----" "---- S3 items owned" "phase X: S3(s-1)" "// store the completed entries" "// the box of this step has landed" "phase Y: S2(s)" "typedef CUresult" > $O/rows_stall_stages.txt 2>&1
cap geom k_geom3d 2 1
cap elast k_rows3d 6 2 --workload elasticity
cap fcm k_assemble_elemset 1 1 --workload fcm
python scripts/ncu_stages.py /tmp/fcm.ncu-rep nutils_b200/csrc/assemble_elemset.cu "// P0: local coordinates" "// P1: 1-D values" "// P1b: solution-dependent" "// P2: geometry at the points" "// P3: values and physical" "if (!MMA && P.ev.enabled) {" "// P4: block entries" "// mass-like block (only D" "// linear forms: thread r owns" "// scatter: bisection" > $O/fcm64_stall_stages.txt 2>&1
cap nurbs k_assemble_elemset 1 1 --workload nurbs_p4
python scripts/ncu_stages.py /tmp/nurbs.ncu-rep nutils_b200/csrc/assemble_elemset.cu "// P0: local coordinates" "// P1: 1-D values" "// P1b: solution-dependent" "// P2: geometry at the points" "// P3: values and physical" "if (!MMA && P.ev.enabled) {" "// P4: block entries" "// mass-like block (only D" "// linear forms: thread r owns" "// scatter: bisection" > $O/nurbs_stall_stages.txt 2>&1
cap rows_p3 k_rows3d 2 1 --degree 3 --n 96
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_poisson.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/launches_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_elasticity.csv python bench.py --steps 2 --warmup 1 --no-cpu --workload elasticity > $O/launches_bench_e.log 2>&1
du -sh $O; ls $O
