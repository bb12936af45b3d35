mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -c 400 gpurun_out/r2_bench_n2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value']); [print(k, v.get('value'), v.get('error')) for k,v in d['extra'].items() if isinstance(v, dict)]"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
