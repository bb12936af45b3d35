mkdir -p gpurun_out
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_diag.json 2> gpurun_out/r2_bench_diag.err
echo "rc=$?"; tail -c 1500 gpurun_out/r2_bench_diag.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_diag.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks']); print({k:(v.get('value'), v.get('error')) if isinstance(v,dict) else v for k,v in (d['extra'] or {}).items()})"
