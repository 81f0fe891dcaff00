# development aid: ncu capture of the fused LR kernel (one B200)
python scripts/ncu_target.py lr > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -k regex:"lr_grad_hess_fano" -o gpurun_out/r2_lr_tlp -f python scripts/ncu_target.py lr > gpurun_out/r2_ncu_lr_tlp.log 2>&1; echo ncu rc=$?
ls -la gpurun_out/r2_lr_tlp.ncu-rep
