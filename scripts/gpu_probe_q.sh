mkdir -p gpurun_out
python -m pytest tests/test_explore_gpu.py tests/test_obstacle_batch_gpu.py tests/test_obstacle_map_gpu.py tests/test_object_map_gpu.py -q 2>&1 | tail -4 > gpurun_out/r2_tests_q.log
for cfg in "--batch 1 --grid 1000" "--batch 32 --grid 1000" "--batch 8 --grid 4000 --ppm 40 --hw 1024 1024"; do python scripts/profile_grid.py $cfg --time 2>&1 | tail -2 >> gpurun_out/r2_grid_times_q.txt; done
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_grid_b32_q.csv python scripts/profile_grid.py --batch 32 --grid 1000 > /dev/null 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_grid_b1_q.csv python scripts/profile_grid.py --batch 1 --grid 1000 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_grid_b32_q.csv gpurun_out/r2_launches_grid_b32_q.md "Value update + ObstacleMapBatch.update, B=32, G=1000, 480x640 (scripts/profile_grid.py)"
python scripts/launch_summary.py gpurun_out/r2_launches_grid_b1_q.csv gpurun_out/r2_launches_grid_b1_q.md "Value update + ObstacleMapBatch.update, B=1, G=1000, 480x640 (scripts/profile_grid.py)"
cat gpurun_out/r2_tests_q.log gpurun_out/r2_grid_times_q.txt; head -24 gpurun_out/r2_launches_grid_b32_q.md
