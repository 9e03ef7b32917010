timeout 900 ncu --set full --clock-control none --import-source on -k regex:"smag_nut_march|explicit3d_march|smag_acc_march" -s 6 -c 3 -o gpurun_out/prof_3d python bench.py --workload TGV512 --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_ncu.log 2>&1
tail -3 gpurun_out/r2s_ncu.log
