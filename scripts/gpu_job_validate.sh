python -m pytest tests -m gpu -x -q > gpurun_out/s6_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/s6_pytest.log
python bench.py > gpurun_out/s6_bench_n1.json 2> gpurun_out/s6_bench_n1.err; echo bench rc=$?; tail -3 gpurun_out/s6_bench_n1.err
for t in lr skinny gemvn; do
  python scripts/ncu_target.py $t > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on --launch-skip 4 -c 1 -k regex:"lr_grad_hess_fano|skinny_dense|gemv_rows_f64" -o gpurun_out/r2_$t -f python scripts/ncu_target.py $t > gpurun_out/r2_ncu_$t.log 2>&1; echo ncu $t rc=$?
done
ls -la gpurun_out/*.ncu-rep | tail -4
