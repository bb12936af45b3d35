mkdir -p gpurun_out
python -m pytest tests/test_blip2_gpu.py -q -s 2>&1 | grep -v Warning > /tmp/full.log; grep -n "outliers=\|passed\|failed\|Error\|error" /tmp/full.log | cut -c1-400 > gpurun_out/r2_blip2_n.log
for v in 1 0; do VLFM_QFORMER_FOLD0=$v python bench.py --steps 20 --warmup 5 --no-extra 2>gpurun_out/r2_bench_n_$v.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('fold0=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['launches_per_step'])" >> gpurun_out/r2_ab_n.txt; done
python scripts/bench_full_step.py --batch 1 --steps 8 --warmup 4 2>/dev/null | tail -1 > gpurun_out/r2_full_step_b1_n.json
cat gpurun_out/r2_blip2_n.log gpurun_out/r2_ab_n.txt; python -c "
import json; d=json.loads(open('gpurun_out/r2_full_step_b1_n.json').read()); print(d['ms_per_step'], d['component_ms_per_step'])"
