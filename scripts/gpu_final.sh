# what the driver runs at round end, on one GPU: GPU tests, smoke, the default bench line, the reference arm
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_K8192.json 2> gpurun_out/final_bench_K8192.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
python - <<PY
import json
for f in ['gpurun_out/final_bench_K8192.json', 'gpurun_out/final_bench_reference.json']:
  d=json.loads([l for l in open(f) if l.startswith('{')][-1])
  print(f, 'value', round(d['value'],3), 'ms', d.get('ms_per_step'), 'e2e', d['e2e']['value'], 'roofline', (d.get('roofline') or {}).get('frac'), 'step frac', (d.get('roofline_step') or {}).get('frac'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'launches', d.get('gpu_launches'), d.get('clocks'))
PY
