import sys, os, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = subprocess.run([sys.executable, 'bench.py', '--workload', sys.argv[1] if len(sys.argv) > 1 else 'TGV256', '--steps', '5', '--warmup', '3', '--no-cpu-baseline'], capture_output=True, text=True).stdout.strip().splitlines()[-1]
d = json.loads(out)
print(os.environ.get('CFD_B200_LIB', 'default').split('/')[-1], round(d['value'], 2), 'Gcell/s', round(d['ms_per_step'], 3), 'ms', {k: round(v, 3) for k, v in d['roofline_step']['kernel_ms'].items()})
