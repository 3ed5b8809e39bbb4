set -x
cd $GRAFT_REPO_ROOT
P="python tools/perf_probe.py --p 8 16 16 --reps 1"
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:syrk_kernel -c 1 -o gpurun_out/r02_k1a_syrk $P > gpurun_out/ev1.log 2>&1
$NCU -k regex:prior_rows_kernel -c 1 -o gpurun_out/r02_k1c_rows_wide $P > gpurun_out/ev2.log 2>&1
$NCU -k regex:potrf_diag_kernel -s 40 -c 1 -o gpurun_out/r02_k5_potrf_diag $P > gpurun_out/ev3.log 2>&1
$NCU -k regex:syrk_update_kernel -s 20 -c 1 -o gpurun_out/r02_k5_syrk_update $P > gpurun_out/ev4.log 2>&1
$NCU -k regex:kyinv_kernel -c 1 -o gpurun_out/r02_k5_kyinv $P > gpurun_out/ev5.log 2>&1
B="python bench.py --steps 1 --warmup 1 --no-full-config --no-small-configs --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_launch_list.csv $B > gpurun_out/ev6.log 2>&1
ls gpurun_out/*.ncu-rep | wc -l
