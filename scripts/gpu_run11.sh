# cluster kernel: timing after the barrier changes + one ncu capture of it (plain layout)
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "paired_spectrum or (matches_oracle and 32768)" 2>&1 | tail -3
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --workload K32768 --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2j_$name.json 2> gpurun_out/r2j_$name.err
  python - <<PY
import json
try:
  d=json.loads([l for l in open('gpurun_out/r2j_$name.json') if l.startswith('{')][-1])
  print('$name', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), {k:round(v,3) for k,v in d['kernel_ms_rank0'].items()})
except Exception as e:
  print('$name FAILED', e); print(open('gpurun_out/r2j_$name.err').read()[-1500:])
PY
}
run plain_cluster CFD_SLAB_SHAPE=32768x8192
run paired_cluster CFD_SLAB_SHAPE=32768x16384
CFD_SLAB_SHAPE=32768x2048 timeout 600 ncu --set full --clock-control none --import-source on -k regex:xlines15 -s 2 -c 1 -o gpurun_out/prof_x15 python bench.py --workload K32768 --gpus 1 --steps 3 --warmup 3 > gpurun_out/r2j_ncu.log 2>&1
tail -3 gpurun_out/r2j_ncu.log
