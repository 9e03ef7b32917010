python scripts/state_hash.py 2>&1 | tail -1
CFD_FFT3_ROWS=8 python scripts/state_hash.py 2>&1 | tail -1
run() { # name, workload, steps, warmup, env...
  name=$1; wl=$2; st=$3; wu=$4; shift; shift; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --gpus 1 --steps $st --warmup $wu --no-cpu-baseline > gpurun_out/r3d_$name.json 2> gpurun_out/r3d_$name.err
  python - <<PY
import json
try:
  d=json.loads([l for l in open('gpurun_out/r3d_$name.json') if l.startswith('{')][-1])
  k = d.get('kernel_ms_rank0') or d.get('roofline_step',{}).get('kernel_ms')
  print('$name', 'ms/step', round(d['ms_per_step'],5), 'value', round(d['value'],2), {a:round(b,4) for a,b in k.items()})
except Exception as e:
  print('$name FAILED', e); print(open('gpurun_out/r3d_$name.err').read()[-1500:])
PY
}
run tgv512_rows8 TGV512 10 3 CFD_FFT3_ROWS=8
run tgv512_default TGV512 10 3 A=1
run tgv256 TGV256 20 3 A=1
run tgv256_r32 TGV256 20 3 CFD_FFT3_ROWS=32
