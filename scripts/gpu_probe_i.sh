mkdir -p gpurun_out
O=gpurun_out/r2_probe_seed2000.txt
python scripts/probe_map_graph.py --batch 32 --seed0 2000 --steps 7 > $O 2>&1
echo "== G=2000 bound 30" >> $O
python scripts/probe_map_graph.py --batch 32 --seed0 2000 --steps 5 --grid 2000 --bound 30 >> $O 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_probe_seed2000_launches.csv python scripts/probe_map_graph.py --batch 32 --seed0 2000 --steps 6 --profile-step 5 > /dev/null 2>&1
cat $O
python scripts/launch_summary.py gpurun_out/r2_probe_seed2000_launches.csv gpurun_out/r2_probe_seed2000_launches.md "seed 2000 step 5" ; head -30 gpurun_out/r2_probe_seed2000_launches.md
