mkdir -p gpurun_out
python -m pytest tests/test_grounding_dino_gpu.py -q -s 2>&1 | grep -v Warning > /tmp/full.log; grep -n "decision test\|passed\|failed\|Error\|batch-1 Ground" /tmp/full.log | cut -c1-700 > gpurun_out/r2_gdino_u.log; tail -c 1200 /tmp/full.log >> gpurun_out/r2_gdino_u.log
O=gpurun_out/r2_fullstep_u.txt
pick='import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(round(d["env_steps_per_s"],1), round(d["ms_per_step"],2), {k: round(v,2) for k,v in d["component_ms_per_step"].items()})'
for B in 1 32; do echo "== B=$B serial" >> $O; VLFM_FULLSTEP_SERIAL=1 python scripts/bench_full_step.py --batch $B --steps 8 --warmup 4 2>/dev/null | python -c "$pick" >> $O; echo "== B=$B streams" >> $O; python scripts/bench_full_step.py --batch $B --steps 8 --warmup 4 2>/dev/null | python -c "$pick" >> $O; done
cat gpurun_out/r2_gdino_u.log $O
