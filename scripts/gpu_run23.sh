N=2; TAG=r2v
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m | head -8
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
run default_again A=1
run c32 CFD_DIST_COPY_CTAS=32
run b4 CFD_DIST_BLOCKS=4
run b4p4 CFD_DIST_BLOCKS=4 CFD_DIST_STENCIL_PARTS=4
run pull CFD_DIST_MODE=pull
