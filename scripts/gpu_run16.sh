run() { # name, workload, env...
  name=$1; wl=$2; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2o_$name.json 2> gpurun_out/r2o_$name.err
  python - <<PY
import json
try:
  d=json.loads([l for l in open('gpurun_out/r2o_$name.json') if l.startswith('{')][-1])
  k = d.get('kernel_ms_rank0') or d.get('roofline_step',{}).get('kernel_ms')
  print('$name', 'ms/step', round(d['ms_per_step'],5), 'value', round(d['value'],2), {a:round(b,4) for a,b in k.items()})
except Exception as e:
  print('$name FAILED', e); print(open('gpurun_out/r2o_$name.err').read()[-1500:])
PY
}
run k256_default K256 A=1
run k256_tx8 K256 CFD_EXPLICIT_TX=8
run k256_tx4 K256 CFD_EXPLICIT_TX=4
run k256_tx2 K256 CFD_EXPLICIT_TX=2
run d2048_default D2048 A=1
run d2048_tx8 D2048 CFD_EXPLICIT_TX=8
run d2048_tx32 D2048 CFD_EXPLICIT_TX=32
run e1024 E1024 A=1
run e1024_c4 E1024 CFD_EXPLICIT_COLS=4
