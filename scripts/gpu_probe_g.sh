mkdir -p gpurun_out
O=gpurun_out/r2_probe_fullstep.txt
pick='import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(round(d["ms_per_step"],2), {k: round(v,2) for k,v in d["component_ms_per_step"].items()})'
echo "== default" >> $O; python scripts/bench_full_step.py --batch 32 2>/dev/null | python -c "$pick" >> $O
echo "== serial" >> $O; VLFM_FULLSTEP_SERIAL=1 python scripts/bench_full_step.py --batch 32 2>/dev/null | python -c "$pick" >> $O
echo "== map graph off" >> $O; VLFM_MAP_GRAPH=0 python scripts/bench_full_step.py --batch 32 2>/dev/null | python -c "$pick" >> $O
echo "== no gdino" >> $O; python scripts/bench_full_step.py --batch 32 --no-gdino 2>/dev/null | python -c "$pick" >> $O
echo "== no gdino, map graph off" >> $O; VLFM_MAP_GRAPH=0 python scripts/bench_full_step.py --batch 32 --no-gdino 2>/dev/null | python -c "$pick" >> $O
python -m pytest tests/test_blip2_gpu.py::test_full_size_vitg_vs_oracle tests/test_grounding_dino_gpu.py::test_detection_decisions_match_the_fp32_twin -q -s 2>&1 | grep -v Warning | tail -25 > gpurun_out/r2_tests_g.log
cat $O; grep -n "outliers=\|decision test\|passed\|failed" gpurun_out/r2_tests_g.log
