#!/bin/bash
# One GPU-box session of round 2: parity tests, the bench line, grid-path timings and ncu evidence.  Run under gpurun from the repo root.
# Only SMALL files go to gpurun_out/ (it is copied back only when <= 64 MiB): ncu reports stay in /tmp, their CSV pages are exported.
mkdir -p gpurun_out
T=${1:-a}
WHAT=${2:-all}
if [[ $WHAT == all || $WHAT == tests ]]; then
python -m pytest tests -m gpu -q -s --deselect tests/test_blip2_gpu.py::test_full_size_vitg_vs_oracle 2>&1 | tail -400 > gpurun_out/r2_gpu_tests_$T.log
python -m pytest tests/test_blip2_gpu.py::test_full_size_vitg_vs_oracle -q -s 2>&1 | tail -40 > gpurun_out/r2_gpu_tests_vitg_$T.log
fi
if [[ $WHAT == gdino ]]; then
python -m pytest tests/test_grounding_dino_gpu.py -q -s 2>&1 | tail -500 > gpurun_out/r2_gdino_tests_$T.log
fi
if [[ $WHAT == all || $WHAT == bench ]]; then
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_$T.json 2> gpurun_out/r2_bench_$T.err
fi
if [[ $WHAT == ab || $WHAT == quick ]]; then
# same box, back to back: deterministic split-K (default) against the atomic one
for v in 1 0 1 0; do
  VLFM_DET_SPLITK=$v python bench.py --steps 20 --warmup 5 --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('det_splitk=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])" >> gpurun_out/r2_ab_$T.txt
done
fi
if [[ $WHAT == quick ]]; then
python -m pytest tests -m gpu -q -s --deselect tests/test_blip2_gpu.py::test_full_size_vitg_vs_oracle 2>&1 | tail -400 > gpurun_out/r2_gpu_tests_$T.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_$T.json 2> gpurun_out/r2_bench_$T.err
for cfg in "--batch 1 --grid 1000" "--batch 32 --grid 1000" "--batch 8 --grid 4000 --ppm 40 --hw 1024 1024"; do
  python scripts/profile_grid.py $cfg --time 2>&1 | tail -2 >> gpurun_out/r2_grid_times_$T.txt
done
fi
if [[ $WHAT == all || $WHAT == prof ]]; then
for cfg in "--batch 1 --grid 1000" "--batch 32 --grid 1000" "--batch 32 --grid 2000" "--batch 8 --grid 4000 --ppm 40 --hw 1024 1024"; do
  python scripts/profile_grid.py $cfg --time 2>&1 | tail -2 >> gpurun_out/r2_grid_times_$T.txt
done
ncu --profile-from-start off --set full --clock-control none -o /tmp/r2_grid_b32 -f python scripts/profile_grid.py --batch 32 --grid 1000 > gpurun_out/r2_ncu_grid_$T.log 2>&1
ncu -i /tmp/r2_grid_b32.ncu-rep --page raw --csv > gpurun_out/r2_ncu_grid_b32_raw_$T.csv 2>/dev/null
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_grid_b1_$T.csv python scripts/profile_grid.py --batch 1 --grid 1000 > /dev/null 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_grid_b32_$T.csv python scripts/profile_grid.py --batch 32 --grid 1000 > /dev/null 2>&1
fi
ls -la gpurun_out | tail -12; du -sh gpurun_out
tail -6 gpurun_out/r2_gpu_tests_$T.log 2>/dev/null; tail -3 gpurun_out/r2_gpu_tests_vitg_$T.log 2>/dev/null; tail -c 300 gpurun_out/r2_bench_$T.err 2>/dev/null; cat gpurun_out/r2_ab_$T.txt 2>/dev/null; cat gpurun_out/r2_grid_times_$T.txt 2>/dev/null
