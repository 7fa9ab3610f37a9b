"""Summarise an ncu `--page source --print-source cuda,sass --csv` dump by source line (top N).
usage: python scripts/ncu_lines.py dump.csv [N]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 45
cur_file = None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
tot = [0, 0, 0]
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if len(r) > 8 and r[0].isdigit():
        k = (cur_file, int(r[0]))
        try:
            s, ie, te = int(r[6] or 0), int(r[7] or 0), int(r[8] or 0)
        except ValueError:
            continue
        a = agg[k]; a[0] += s; a[1] += ie; a[2] += te; a[3] = r[1].strip()[:90]
        tot[0] += s; tot[1] += ie; tot[2] += te
print("total samples", tot[0], "warp-inst", tot[1], "thread-inst", tot[2], "avg active", tot[2] / max(tot[1], 1))
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:N]:
    print(f"{100*a[0]/tot[0]:5.1f}% smp {100*a[1]/tot[1]:5.1f}% inst act {a[2]/max(a[1],1):5.1f}  {k[0]}:{k[1]:<4d} {a[3]}")
