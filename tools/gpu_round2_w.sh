#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/w_launch_train.csv python train_bench.py --model upconv --steps 2 --warmup 1 --no-graph > gpurun_out/w_ncu_train.log 2>&1
python - <<'PY'
import csv,re
from collections import OrderedDict
rows=[r for r in csv.reader(open('gpurun_out/w_launch_train.csv',errors='replace')) if len(r)>14 and r[0].isdigit()]
# last step only: take the last third of launches
n=len(rows)//3
rows=rows[-n:]
agg=OrderedDict()
for r in rows:
    name=re.sub(r"\(.*","",r[4]).replace("void ","").strip()[-70:]
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=float(r[-1].replace(",",""))
tot=sum(a[1] for a in agg.values())
print("launches in one step: %d, total kernel time %.2f ms (cold, serialised)"%(len(rows),tot/1e6))
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:28]:
    print("%6.1f us  %4.1f%%  x%-3d %s"%(t/1e3,100*t/tot,c,k))
PY
