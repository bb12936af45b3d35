mkdir -p gpurun_out
python -m pytest tests/test_x2_gpu.py -q -x 2>&1 | tail -25 > gpurun_out/r2_x2_tests.log
python -m pytest tests/test_blip2_gpu.py -q -s 2>&1 | grep -v Warning > /tmp/full.log; grep -n "outliers=\|passed\|failed\|Error\|error" /tmp/full.log | cut -c1-400 > gpurun_out/r2_blip2_j.log; tail -c 1500 /tmp/full.log >> gpurun_out/r2_blip2_j.log
for v in 1 0; do VLFM_QFORMER_X2=$v python bench.py --steps 20 --warmup 5 --no-extra 2>gpurun_out/r2_bench_j_$v.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('qformer_x2=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['launches_per_step'])" >> gpurun_out/r2_ab_j.txt; done
cat gpurun_out/r2_x2_tests.log | tail -12; cat gpurun_out/r2_blip2_j.log | head -20; cat gpurun_out/r2_ab_j.txt; tail -3 gpurun_out/r2_bench_j_1.err
