mkdir -p gpurun_out
python -m pytest tests/test_x2_gpu.py tests/test_gemm_gpu.py -q 2>&1 | tail -8 > gpurun_out/r2_x2_tests_l.log
python -m pytest tests/test_blip2_gpu.py -q -s 2>&1 | grep -v Warning > /tmp/full.log; grep -n "outliers=\|passed\|failed\|Error\|error" /tmp/full.log | cut -c1-400 > gpurun_out/r2_blip2_l.log
for v in 1 0; do VLFM_QFORMER_X2=$v python bench.py --steps 20 --warmup 5 --no-extra 2>gpurun_out/r2_bench_l_$v.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('qformer_x2=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['launches_per_step'])" >> gpurun_out/r2_ab_l.txt; done
VLFM_NO_GRAPH=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_step_x2.csv python scripts/profile_step.py > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_step_x2.csv gpurun_out/r2_launches_step_x2.md "configs[1] step, batch 1, un-graphed, x2 Q-Former"
python __graft_entry__.py smoke 2>&1 | tail -2 > gpurun_out/r2_smoke_l.log
cat gpurun_out/r2_x2_tests_l.log gpurun_out/r2_blip2_l.log gpurun_out/r2_ab_l.txt gpurun_out/r2_smoke_l.log; head -30 gpurun_out/r2_launches_step_x2.md; tail -3 gpurun_out/r2_bench_l_1.err
