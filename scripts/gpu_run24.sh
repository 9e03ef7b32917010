N=4; TAG=r2w
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { # name, env...
  name=$1; shift
  env "$@" CFD_SKIP_K32768=1 timeout 900 $TR --master-port 29520 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_n${N}_$name.json 2> gpurun_out/${TAG}_n${N}_$name.err
  python - <<PY
import json
try:
  d=json.loads([l for l in open('gpurun_out/${TAG}_n${N}_$name.json') if l.startswith('{')][-1])
  print('$name', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), {k:round(v,3) for k,v in d['kernel_ms_rank0'].items()})
except Exception as e:
  print('$name FAILED', e); print(open('gpurun_out/${TAG}_n${N}_$name.err').read()[-1500:])
PY
}
run default A=1
run c12 CFD_DIST_COPY_CTAS=12
run c24 CFD_DIST_COPY_CTAS=24
