mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'colsum|sum_parts|rankb' --csv --log-file gpurun_out/r02_launches_c4_vtlpm.csv \
    python bench.py --workload c4 --rules VTL-PM --steps 2 --warmup 1 > gpurun_out/r02_ncu_c4.log 2>&1
echo "c4 launch list rc=$?"
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02_launches_c4_vtlpm.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=i;break
hdr=rows[h]; kn=hdr.index('Kernel Name'); mn=hdr.index('Metric Name'); mv=hdr.index('Metric Value'); mu=hdr.index('Metric Unit'); gi=hdr.index('Grid Size')
agg=collections.OrderedDict()
for r in rows[h+1:]:
    if len(r)<=mv: continue
    key=(r[kn].split('(')[0], r[gi], r[mn], r[mu])
    a=agg.setdefault(key,[0,0.0]); a[0]+=1; a[1]+=float(r[mv].replace(',',''))
for k,a in agg.items():
    print(k[0][:70], k[1], k[2], k[3], a[0], round(a[1]/a[0],4))
PY
