# round-2 evidence for the headline workload: launch list + full capture of one chained step's kernels
# (reports are turned into CSV on the box: gpurun_out/ may only bring back 64 MiB)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_chain.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"explicit2d|rfft_rows|xlines_kernel|irfft_rows|correct2d" -s 8 -c 5 -o /tmp/prof_r02_chain python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_f.log 2>&1
ncu -i /tmp/prof_r02_chain.ncu-rep --page raw --csv > gpurun_out/r02_chain_raw.csv 2>/dev/null
ls -la /tmp/prof_r02_chain.ncu-rep
timeout 900 ncu --set full --clock-control none -k regex:"explicit2d" -s 2 -c 1 -o /tmp/prof_r02_e1024 python bench.py --workload E1024 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_e.log 2>&1
ncu -i /tmp/prof_r02_e1024.ncu-rep --page raw --csv > gpurun_out/r02_e1024_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:"smag_nut_march|explicit3d_march|smag_acc_march|xlines3|rfft_rows3|divergence3d|correct3d" -s 11 -c 11 -o /tmp/prof_r02_tgv python bench.py --workload TGV512 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_t.log 2>&1
ncu -i /tmp/prof_r02_tgv.ncu-rep --page raw --csv > gpurun_out/r02_tgv512_raw.csv 2>/dev/null
du -sh gpurun_out
