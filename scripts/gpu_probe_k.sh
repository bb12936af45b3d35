mkdir -p gpurun_out
VLFM_NO_GRAPH=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_step_x2.csv python scripts/profile_step.py > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_step_x2.csv gpurun_out/r2_launches_step_x2.md "configs[1] step, batch 1, un-graphed, x2 Q-Former"
head -40 gpurun_out/r2_launches_step_x2.md
grep "gemm_f16x2\|attention_f32" gpurun_out/r2_launches_step_x2.csv | awk -F'","' '{print $5, $9, $NF}' | sed 's/"//g' | sort | uniq -c | sort -k1 -n -r | head -60
