N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"

timeout 300 $TR --master-port 29512 tests/mgpu_worker.py 128 128 64 2 2>&1 | grep -E "MGPU|rror" | tail -2
timeout 900 $TR --master-port 29520 bench.py --gpus $N --steps 20 --warmup 5 --workload TGV512 > gpurun_out/r3b_n${N}_tgv512.json 2> gpurun_out/r3b_n${N}_tgv512.err
python - <<PY
import json
try:
  d=json.loads([l for l in open('gpurun_out/r3b_n${N}_tgv512.json') if l.startswith('{')][-1])
  print('tgv512', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],2), 'parity', (d.get('parity') or {}).get('bitwise'), {k:round(v,3) for k,v in d['kernel_ms_rank0'].items()})
except Exception as e:
  print('FAILED', e); print(open('gpurun_out/r3b_n${N}_tgv512.err').read()[-1500:])
PY
