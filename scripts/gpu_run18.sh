python -m pytest tests -m gpu -q -x -k "batched or E1024 or ensemble or members" 2>&1 | tail -4
run() { # name, workload, steps, warmup, env...
  name=$1; wl=$2; st=$3; wu=$4; shift; shift; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --gpus 1 --steps $st --warmup $wu --no-cpu-baseline > gpurun_out/r2q_$name.json 2> gpurun_out/r2q_$name.err
  python - <<PY
import json
try:
  d=json.loads([l for l in open('gpurun_out/r2q_$name.json') if l.startswith('{')][-1])
  k = d.get('kernel_ms_rank0') or d.get('roofline_step',{}).get('kernel_ms')
  print('$name', 'ms/step', round(d['ms_per_step'],5), 'value', round(d['value'],2), {a:round(b,4) for a,b in k.items()})
except Exception as e:
  print('$name FAILED', e); print(open('gpurun_out/r2q_$name.err').read()[-1500:])
PY
}
run e1024_wrap E1024 20 5 A=1
run e1024_wrap_tx64 E1024 20 5 CFD_EXPLICIT_TX=64
run e1024_wrap_tx32 E1024 20 5 CFD_EXPLICIT_TX=32
run e1024_nowrap E1024 20 5 CFD_EXPLICIT_WRAP=0
