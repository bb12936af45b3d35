mkdir -p gpurun_out
B=1 VLFM_GDINO_GRAPH=0 python scripts/profile_gdino.py 2>&1 | grep -v Warning | cut -c1-230 > gpurun_out/r2_gdino_profile_b1.txt
head -16 gpurun_out/r2_gdino_profile_b1.txt
python - <<'PY'
import re
rows=[]
for l in open('gpurun_out/r2_gdino_profile_b1.txt'):
    m=re.match(r'\s*(.+?)\s{2,}[\d.]+%.*?\s+([\d.]+)(us|ms)\s+[\d.]+%\s+([\d.]+)(us|ms)\s+([\d.]+)(us|ms)\s+(\d+)\s*$', l)
    if m: rows.append((m.group(1)[:70], m.group(4)+m.group(5), int(m.group(8))))
for r in rows[:45]: print(r)
PY
