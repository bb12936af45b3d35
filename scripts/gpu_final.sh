mkdir -p gpurun_out
python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2>/dev/null
python __graft_entry__.py smoke 2>&1 | tail -1
tail -c 300 gpurun_out/r2_bench_final.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['launches_per_step'], d['roofline']['not_replayed']); [print(k, v.get('value'), v.get('component_ms_per_step'), v.get('error')) for k,v in d['extra'].items() if isinstance(v, dict)]"
cut -c1-400 gpurun_out/r2_bench_reference.json
