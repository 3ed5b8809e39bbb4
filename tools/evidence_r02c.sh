# Round-2 evidence of the final build (K3 on the tensor pipe, K2 / K1c-finalize rework): GPU test log with the PARITY lines, the
# default bench line, launch lists of one bench step and of post-intervention trials, captures of the new / changed kernels.
set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q -s > gpurun_out/r02c_pytest_gpu_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r02c_pytest_gpu_parity.log
python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
B="python bench.py --steps 1 --warmup 1 --no-full-config --no-small-configs --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02c_bench_launch_list.csv $B > gpurun_out/evc1.log 2>&1
R="python tools/refresh_probe.py --n-obs 10000 --p 100 100 100 --trials 6"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02c_refresh_launches.csv $R > gpurun_out/evc2.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
R2="python tools/refresh_probe.py --n-obs 2000 --p 100 100 100 --trials 6"
$NCU -k regex:posterior_fit_kernel -s 3 -c 1 -o gpurun_out/r02c_k2_fit $R2 > gpurun_out/evc3.log 2>&1
$NCU -k regex:sweep_mma_kernel -s 3 -c 1 -o gpurun_out/r02c_k3_mma_refresh $R2 > gpurun_out/evc4.log 2>&1
python tools/refresh_probe.py --n-obs 10000 --p 100 100 100 --trials 12 > gpurun_out/r02c_refresh_probe.json 2>/dev/null
tail -2 gpurun_out/r02c_pytest_gpu_parity.log
