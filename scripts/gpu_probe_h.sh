mkdir -p gpurun_out
python -m pytest tests/test_blip2_gpu.py::test_full_size_vitg_vs_oracle tests/test_grounding_dino_gpu.py::test_detection_decisions_match_the_fp32_twin -q -s 2>&1 | grep -v Warning > gpurun_out/r2_tests_h_full.log
grep -n "outliers=\|decision test\|passed\|failed\|Error" gpurun_out/r2_tests_h_full.log | cut -c1-900 > gpurun_out/r2_tests_h.log
python -m pytest tests/test_explore_gpu.py tests/test_obstacle_batch_gpu.py tests/test_obstacle_map_gpu.py -q 2>&1 | tail -5 >> gpurun_out/r2_tests_h.log
for cfg in "--batch 1 --grid 1000" "--batch 32 --grid 1000"; do python scripts/profile_grid.py $cfg --time 2>&1 | tail -2 >> gpurun_out/r2_grid_times_h.txt; done
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_h.json 2> gpurun_out/r2_bench_h.err
tail -c 600 gpurun_out/r2_tests_h_full.log > gpurun_out/r2_tests_h_tail.log; rm gpurun_out/r2_tests_h_full.log
cat gpurun_out/r2_tests_h.log gpurun_out/r2_grid_times_h.txt; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_h.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac']); [print(k, v.get('value'), v.get('component_ms_per_step'), v.get('error')) for k,v in d['extra'].items() if isinstance(v, dict)]"
