mkdir -p gpurun_out
VLFM_NO_GRAPH=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'gemm_f16x2|attention_f32|split_x2' -c 30 -o /tmp/r2_x2 -f python scripts/profile_step.py > gpurun_out/r2_ncu_x2.log 2>&1
ncu -i /tmp/r2_x2.ncu-rep --page raw --csv > gpurun_out/r2_ncu_x2_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_ncu_x2_raw.csv gpurun_out/r2_ncu_x2 2>&1 | tail -3
VLFM_NO_GRAPH=1 ncu --profile-from-start off --set full --clock-control none -k regex:'layernorm_reduce' -c 6 -o /tmp/r2_ln -f python scripts/profile_step.py > /dev/null 2>&1
ncu -i /tmp/r2_ln.ncu-rep --page raw --csv > gpurun_out/r2_ncu_ln_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_ncu_ln_raw.csv gpurun_out/r2_ncu_ln 2>&1 | tail -3
python -m pytest tests/test_gemm_gpu.py tests/test_x2_gpu.py tests/test_attention_ln_gpu.py -q 2>&1 | tail -3
python -m pytest tests/test_blip2_gpu.py -q -s 2>&1 | grep "outliers=\|passed\|failed"
python bench.py --steps 20 --warmup 5 --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
head -34 gpurun_out/r2_ncu_x2.md; cat gpurun_out/r2_ncu_ln.md
