#!/bin/bash
# One GPU-box session of round 2: parity tests, the bench line, grid-path timings and ncu evidence.  Run under gpurun from the repo root.
mkdir -p gpurun_out
T=${1:-a}
python -m pytest tests -m gpu -q -s --deselect tests/test_blip2_gpu.py::test_full_size_vitg_vs_oracle 2>&1 | tail -120 > gpurun_out/r2_gpu_tests_$T.log
python -m pytest tests/test_blip2_gpu.py::test_full_size_vitg_vs_oracle -q -s 2>&1 | tail -40 > gpurun_out/r2_gpu_tests_vitg_$T.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_$T.json 2> gpurun_out/r2_bench_$T.err
for cfg in "--batch 1 --grid 1000" "--batch 32 --grid 1000" "--batch 32 --grid 2000" "--batch 8 --grid 4000 --ppm 40 --hw 1024 1024"; do
  python scripts/profile_grid.py $cfg --time 2>&1 | tail -2 >> gpurun_out/r2_grid_times_$T.txt
done
ncu --profile-from-start off --set full --clock-control none -o gpurun_out/r2_grid_b32_$T -f python scripts/profile_grid.py --batch 32 --grid 1000 > gpurun_out/r2_ncu_grid_$T.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_grid_b1_$T.csv python scripts/profile_grid.py --batch 1 --grid 1000 > /dev/null 2>&1
tail -5 gpurun_out/r2_gpu_tests_$T.log; tail -3 gpurun_out/r2_gpu_tests_vitg_$T.log; tail -c 400 gpurun_out/r2_bench_$T.err; cat gpurun_out/r2_grid_times_$T.txt
