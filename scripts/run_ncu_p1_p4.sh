NCU="ncu --set full --clock-control none"
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
O=gpurun_out/ncu
mkdir -p $O gpurun_out/bench
$NCU -k regex:k_rows3d --launch-skip 2 -c 1 -o /tmp/rows_p1 $B --degree 1 > $O/rows_p1.log 2>&1
ncu -i /tmp/rows_p1.ncu-rep --page raw --csv > $O/rows_p1_raw.csv 2>/dev/null; ncu -i /tmp/rows_p1.ncu-rep --page details > $O/rows_p1_details.txt 2>/dev/null
$NCU -k regex:k_rows3d --launch-skip 2 -c 1 -o /tmp/rows_p4 $B --degree 4 --n 48 > $O/rows_p4.log 2>&1
ncu -i /tmp/rows_p4.ncu-rep --page raw --csv > $O/rows_p4_raw.csv 2>/dev/null; ncu -i /tmp/rows_p4.ncu-rep --page details > $O/rows_p4_details.txt 2>/dev/null
grep -E "Duration|Registers Per" $O/rows_p1_details.txt $O/rows_p4_details.txt
