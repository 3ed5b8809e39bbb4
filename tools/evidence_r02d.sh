# Captures of the kernels rebuilt on the tensor-map TMA pipeline (K1a SYRK, K5 trailing update) and of K5's blocked
# diagonal-block kernel, plus smoke() of the final build.  (--kernel-name-base demangled: the filter sees the template arguments)
set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02d_smoke.log 2>&1
P="python tools/perf_probe.py --p 8 16 16 --reps 1"
NCU="ncu --set full --clock-control none --import-source on -f --kernel-name-base demangled"
timeout 150 $NCU -k regex:PriorSyrkPlan -c 1 -o gpurun_out/r02d_k1a_syrk_tma $P > gpurun_out/evd1.log 2>&1
timeout 150 $NCU -k "regex:tma_tile_kernel<cbo::SyrkPlan" -s 2 -c 1 -o gpurun_out/r02d_k5_trailing_tma $P > gpurun_out/evd2.log 2>&1
timeout 150 $NCU -k regex:potrf_diag_kernel -s 40 -c 1 -o gpurun_out/r02d_k5_potrf_diag $P > gpurun_out/evd3.log 2>&1
tail -2 gpurun_out/r02d_smoke.log
