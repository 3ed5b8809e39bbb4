set -x
cd $GRAFT_REPO_ROOT
P="python tools/perf_probe.py --p 8 16 16 --reps 1"
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:syrk_kernel -c 1 -o gpurun_out/r02_k1a_syrk $P > gpurun_out/ev1.log 2>&1
$NCU -k regex:prior_rows_kernel -c 1 -o gpurun_out/r02_k1c_rows_wide $P > gpurun_out/ev2.log 2>&1
$NCU -k regex:potrf_diag_kernel -s 40 -c 1 -o gpurun_out/r02_k5_potrf_diag $P > gpurun_out/ev3.log 2>&1
$NCU -k regex:syrk_update_kernel -s 20 -c 1 -o gpurun_out/r02_k5_syrk_update $P > gpurun_out/ev4.log 2>&1
$NCU -k regex:kyinv_kernel -c 1 -o gpurun_out/r02_k5_kyinv $P > gpurun_out/ev5.log 2>&1
G="python tools/golden_probe.py simplified_coral"
$NCU -k regex:prior_pair_kernel -s 1 -c 1 -o gpurun_out/r02_k1p_pair $G > gpurun_out/ev6.log 2>&1
$NCU -k regex:sweep_kernel -s 1 -c 1 -o gpurun_out/r02_k3_sweep $G > gpurun_out/ev7.log 2>&1
$NCU -k regex:pair_tables_kernel -s 1 -c 1 -o gpurun_out/r02_k1p_tables $G > gpurun_out/ev8.log 2>&1
R="python tools/refresh_probe.py"
$NCU -k regex:prior_rows_kernel -s 1 -c 1 -o gpurun_out/r02_k1c_rows_narrow $R > gpurun_out/ev9.log 2>&1
$NCU -k regex:sweep_kernel -s 3 -c 1 -o gpurun_out/r02_k3_sweep_refresh $R > gpurun_out/ev10.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_refresh_launches.csv $R > gpurun_out/ev11.log 2>&1
python tools/refresh_probe.py --p 100 100 100 > gpurun_out/r02_refresh_probe.json 2>/dev/null
ls -la gpurun_out/*.ncu-rep | wc -l
