mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lower.py tests/test_gpu_guards.py -q -m gpu > gpurun_out/r2_t14.log 2>&1; echo "tests rc=$?"
grep -E "^(FAILED|ERROR|E  )|passed|failed" gpurun_out/r2_t14.log | head -20
for V in 4 1; do
TAL_ROWS_DENSE_VPT=$V timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'colsum' --csv --log-file gpurun_out/r02_launches_c4_vpt$V.csv \
    python bench.py --workload c4 --rules VTL-PM --steps 2 --warmup 1 > gpurun_out/r02_ncu_c4.log 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02_launches_c4_vpt$V.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=i;break
hdr=rows[h]; kn=hdr.index('Kernel Name'); mn=hdr.index('Metric Name'); mv=hdr.index('Metric Value'); mu=hdr.index('Metric Unit'); gi=hdr.index('Grid Size')
agg=collections.OrderedDict()
for r in rows[h+1:]:
    if len(r)<=mv: continue
    key=(r[kn].split('(')[0], r[gi], r[mn], r[mu])
    a=agg.setdefault(key,[0,0.0]); a[0]+=1; a[1]+=float(r[mv].replace(',',''))
for k,a in agg.items():
    if 'rows_dense' in k[0]: print("VPT=$V", k[0][:70], k[1], k[2], k[3], a[0], round(a[1]/a[0],4))
PY
done
timeout 600 python bench.py --workload c4 --rules VTL-PM --steps 8 --warmup 2 > gpurun_out/r2_bench14_c4.json 2> gpurun_out/r2_bench14_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_bench14_c4.json").read().strip().splitlines()[-1])
    r=d["rules"]["VTL-PM"]; print("c4 VTL-PM", r["value"], r["ms_per_step"], r["roofline"]["ms_per_launch"], r["roofline"]["frac"])
except Exception as e: print("parse", e); print(open("gpurun_out/r2_bench14_c4.err").read()[-1200:])
PY
