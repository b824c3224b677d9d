ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_stream|k_exact|k_corr" -s 6 -c 6 --csv --log-file gpurun_out/r02g_k1split.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r02g_k1split.csv') if l.startswith('"'))]
h=rows[0]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value'); iid=h.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault((r[iid], r[ik][:40]),{})[r[im]]=r[iv]
for k,v in cur.items(): print(k, v)
PY
