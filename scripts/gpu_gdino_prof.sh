mkdir -p gpurun_out
python scripts/profile_gdino.py 2>&1 | grep -v Warning | cut -c1-200 > gpurun_out/r2_gdino_profile_b32.txt
head -70 gpurun_out/r2_gdino_profile_b32.txt
