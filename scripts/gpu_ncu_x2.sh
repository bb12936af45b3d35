mkdir -p gpurun_out
VLFM_NO_GRAPH=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'gemm_f16x2|attention_f32|layernorm_reduce|split_x2' -c 40 -o /tmp/r2_x2 -f python scripts/profile_step.py > gpurun_out/r2_ncu_x2.log 2>&1
ncu -i /tmp/r2_x2.ncu-rep --page raw --csv > gpurun_out/r2_ncu_x2_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_ncu_x2_raw.csv gpurun_out/r2_ncu_x2.md 2>&1 | tail -3
compute-sanitizer --tool memcheck python scripts/sanitize_r2.py > gpurun_out/r2_sanitizer_memcheck.txt 2>&1
compute-sanitizer --tool racecheck python scripts/sanitize_r2.py > gpurun_out/r2_sanitizer_racecheck.txt 2>&1
tail -4 gpurun_out/r2_sanitizer_memcheck.txt; tail -4 gpurun_out/r2_sanitizer_racecheck.txt; head -30 gpurun_out/r2_ncu_x2.md; du -sh gpurun_out
