python -m pytest tests -m gpu -q -x 2>&1 | tail -4
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --workload K32768 --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2k_$name.json 2> gpurun_out/r2k_$name.err
  python - <<PY
import json
try:
  d=json.loads([l for l in open('gpurun_out/r2k_$name.json') if l.startswith('{')][-1])
  print('$name', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), {k:round(v,3) for k,v in d['kernel_ms_rank0'].items()})
except Exception as e:
  print('$name FAILED', e); print(open('gpurun_out/r2k_$name.err').read()[-1500:])
PY
}
run plain_cluster CFD_SLAB_SHAPE=32768x8192
run plain_split CFD_SLAB_SHAPE=32768x8192 CFD_X15=split
