# development aid: ncu captures of the round-2 streaming kernels (one B200)
for t in syrk tall; do
  python scripts/ncu_target.py $t > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -k regex:"dsyrk128|tall128" -o gpurun_out/r2_$t -f python scripts/ncu_target.py $t > gpurun_out/r2_ncu_$t.log 2>&1; echo ncu $t rc=$?
done
ls -la gpurun_out/r2_syrk.ncu-rep gpurun_out/r2_tall.ncu-rep
