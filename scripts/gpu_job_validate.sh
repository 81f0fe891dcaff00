# development aid: full GPU test-suite + ncu captures of the round-2 streaming kernels (one B200)
python -m pytest tests -m gpu -x -q > gpurun_out/s7_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/s7_pytest.log
for t in lr skinny gemvn; do
  python scripts/ncu_target.py $t > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -k regex:"lr_grad_hess_fano|skinny_dense|gemv_rows_f64" -o gpurun_out/r2_$t -f python scripts/ncu_target.py $t > gpurun_out/r2_ncu_$t.log 2>&1; echo ncu $t rc=$?
done
ls -la gpurun_out/r2_lr.ncu-rep gpurun_out/r2_skinny.ncu-rep gpurun_out/r2_gemvn.ncu-rep
python scripts/gpu_probe.py lr 2>&1 | grep lr_fused
