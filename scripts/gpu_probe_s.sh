mkdir -p gpurun_out
O=gpurun_out/r2_fullstep_streams.txt
pick='import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(round(d["env_steps_per_s"],1), round(d["ms_per_step"],2), {k: round(v,2) for k,v in d["component_ms_per_step"].items()})'
for B in 32 1; do
echo "== B=$B three streams" >> $O; python scripts/bench_full_step.py --batch $B --steps 8 --warmup 4 2>gpurun_out/err_s.txt | python -c "$pick" >> $O; tail -2 gpurun_out/err_s.txt >> $O
echo "== B=$B serial" >> $O; VLFM_FULLSTEP_SERIAL=1 python scripts/bench_full_step.py --batch $B --steps 8 --warmup 4 2>/dev/null | python -c "$pick" >> $O
done
cat $O
