mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --extra-budget 170 > gpurun_out/r2_bench_diag.json 2> gpurun_out/r2_bench_diag.err
echo "rc=$?"; tail -c 2500 gpurun_out/r2_bench_diag.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_diag.json').read().strip().splitlines()[-1]); print(d['value']); print({k:(v.get('value'), v.get('error')) if isinstance(v,dict) else v for k,v in (d['extra'] or {}).items()})"
