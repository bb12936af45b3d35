mkdir -p gpurun_out
(echo "== graph on"; python scripts/probe_map_graph.py --batch 32; echo "== graph off"; VLFM_MAP_GRAPH=0 python scripts/probe_map_graph.py --batch 32) > gpurun_out/r2_probe_map_graph.txt 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_probe_launches_graph.csv python scripts/probe_map_graph.py --batch 32 --profile-step 8 > /dev/null 2>&1
python -m pytest tests/test_grounding_dino_gpu.py -q -s -k "decisions or head_kernels or own_forward" 2>&1 | tail -30 > gpurun_out/r2_gdino_f.log
for v in 1 0; do VLFM_DET_SPLITK=$v python bench.py --steps 20 --warmup 5 --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('det_splitk=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])" >> gpurun_out/r2_ab_f.txt; done
python -m pytest tests/test_blip2_gpu.py tests/test_gemm_gpu.py -q 2>&1 | tail -5 >> gpurun_out/r2_ab_f.txt
cat gpurun_out/r2_probe_map_graph.txt; cat gpurun_out/r2_ab_f.txt; tail -12 gpurun_out/r2_gdino_f.log
