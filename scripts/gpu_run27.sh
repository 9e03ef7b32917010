run() { # name, workload, steps, warmup, env...
  name=$1; wl=$2; st=$3; wu=$4; shift; shift; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --gpus 1 --steps $st --warmup $wu --no-cpu-baseline > gpurun_out/r2x_$name.json 2> gpurun_out/r2x_$name.err
  python - <<PY
import json
try:
  d=json.loads([l for l in open('gpurun_out/r2x_$name.json') if l.startswith('{')][-1])
  k = d.get('kernel_ms_rank0') or d.get('roofline_step',{}).get('kernel_ms')
  print('$name', 'ms/step', round(d['ms_per_step'],5), 'value', round(d['value'],2), {a:round(b,4) for a,b in k.items()})
except Exception as e:
  print('$name FAILED', e); print(open('gpurun_out/r2x_$name.err').read()[-1500:])
PY
}
run d2048 D2048 200 10 A=1
run d2048_graph D2048 200 10 CFD_GRAPH_CELLS=5000000
run d2048_c2tx32 D2048 200 10 CFD_EXPLICIT_COLS=2 CFD_EXPLICIT_TX=32
run d2048_le5 D2048 200 10 CFD_XLINES_LE=5
run d2048_le5_graph D2048 200 10 CFD_XLINES_LE=5 CFD_GRAPH_CELLS=5000000
run d2048_rows1 D2048 200 10 CFD_FFT_ROWS_SHIFT=1
