# K3 probe: GPU tests, then sweep-stage timings at n_int = 45 / 32 and post-intervention trials, for the shipped library and
# for build variants found under cbo_with_oop_b200/build/lib_*.so (development aid).
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_parity_gpu.py tests/test_edge_gpu.py -m gpu -x -q > gpurun_out/k3_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/k3_pytest.log
run() {
  python tools/perf_probe.py --n-obs 256 --p 100 100 100 --n-int 45 --reps 5 > gpurun_out/k3_$1_n45.json 2> gpurun_out/k3_$1.err
  python tools/perf_probe.py --n-obs 256 --p 100 100 100 --n-int 32 --reps 5 > gpurun_out/k3_$1_n32.json 2>> gpurun_out/k3_$1.err
  python tools/refresh_probe.py --n-obs 2000 --p 100 100 100 --trials 10 > gpurun_out/k3_$1_refresh.json 2>> gpurun_out/k3_$1.err
}
run shipped
ncu --set full --clock-control none --import-source on -k regex:sweep_mma -c 1 -f -o gpurun_out/r02_k3_mma python tools/perf_probe.py --n-obs 256 --p 100 100 100 --n-int 45 --reps 1 > gpurun_out/k3_ncu.log 2>&1
for v in cbo_with_oop_b200/build/lib_*.so; do
  [ -f "$v" ] || continue
  cp "$v" cbo_with_oop_b200/libcbo_b200.so
  run $(basename $v .so)
done
tail -3 gpurun_out/k3_pytest.log
