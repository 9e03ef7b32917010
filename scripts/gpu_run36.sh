timeout 900 ncu --set full --clock-control none -k regex:"smag_nut_march|explicit3d_march|xlines3|rfft_rows3|cfft_lines|divergence3d|correct3d" -s 10 -c 10 -o /tmp/prof_r02_tgv_b python bench.py --workload TGV512 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r3e_t.log 2>&1
ncu -i /tmp/prof_r02_tgv_b.ncu-rep --page raw --csv > gpurun_out/r02_tgv512_fused_raw.csv 2>/dev/null
ls -la gpurun_out/
