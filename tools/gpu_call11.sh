# 1 GPU: per-phase breakdown of the fused step, launch list of one C4 VTL-PM step
mkdir -p gpurun_out
TAL_BREAKDOWN=1 timeout 400 python bench.py --steps 32 --warmup 3 --no-subrecords --no-both --no-cpu > gpurun_out/r02_breakdown_g1.json 2> gpurun_out/r02_breakdown_g1.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_breakdown_g1.json").read().strip().splitlines()[-1])
    print(1, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d.get("breakdown_us"))
except Exception as e:
    print("parse", e); print(open("gpurun_out/r02_breakdown_g1.err").read()[-1500:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c4_vtlpm.csv \
    python bench.py --workload c4 --rules VTL-PM --steps 3 --warmup 1 > gpurun_out/r02_ncu_c4.log 2>&1
echo "c4 launch list rc=$?"
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02_launches_c4_vtlpm.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=i;break
hdr=rows[h]; kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value'); mu=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[h+1:]:
    if len(r)<=mv: continue
    n=r[kn].split('(')[0]; v=float(r[mv].replace(',',''))
    u=r[mu]
    if u=='ns': v/=1e6
    elif u=='us': v/=1e3
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
for n,a in sorted(agg.items(), key=lambda x:-x[1][1])[:16]:
    print(f"{n[:90]:90s} {a[0]:5d} {a[1]:10.3f} ms {a[1]/a[0]:9.4f}")
PY
