#!/bin/bash
# tools/ncu_profile_txt.sh <report.ncu-rep> <cells> "<header line>"  -> the text summary kept under profiles/
rep=$1; cells=$2
echo "# $3"
python tools/ncu_summary.py $rep $cells
echo; echo "SASS opcode histogram (executed warp instructions per 32 cell-updates):"
python tools/ncu_ops.py $rep $cells 22
echo; echo "stall reasons (warps per issue-active cycle):"
ncu -i $rep --page raw --csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
out=[]
for i,n in enumerate(h):
    if 'issue_stalled' in n and 'per_issue_active' in n:
        try: out.append((float(r[i]), n))
        except: pass
for v,n in sorted(out, reverse=True)[:8]: print('  %6.2f %s'%(v,n.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
for i,n in enumerate(h):
    if n in ('sm__cycles_active.avg','sm__cycles_active.max','sm__cycles_active.min'): print('  %s %s' % (n, r[i]))
"
