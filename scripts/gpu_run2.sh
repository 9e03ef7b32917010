set -x
python -m pytest tests -m gpu -q --maxfail=6 -k "not slab_decomposition" 2>&1 | tail -25
for v in default bal0 tw0; do
  case $v in
    default) ENVV="";;
    bal0) ENVV="CFD_XLINES_BAL=0";;
    tw0) ENVV="CFD_B200_LIB=$PWD/jax-cfd_b200/lib/libcfd_b200_tw0.so";;
  esac
  env $ENVV python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_$v.json 2> gpurun_out/r2b_bench_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2b_bench_$v.json'))
print('$v', round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['roofline_step']['kernel_ms'].items()})
PY
done
