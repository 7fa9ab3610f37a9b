"""Aggregate an ncu source-page CSV by enclosing function of each source line (heuristic: the last
line above it that looks like a function definition).  usage: ncu_regions.py report.ncu-rep"""
import collections, csv, io, os, re, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vecchio_b200", "csrc")
fn_of = {}
def load(fname):
    path = os.path.join(root, fname)
    if not os.path.exists(path): return None
    names, cur = {}, "?"
    pat = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:VKD|__device__|__global__)[^;{]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(")
    for i, line in enumerate(open(path), 1):
        m = pat.match(line)
        if m and not line.strip().endswith(";"): cur = m.group(1)
        names[i] = cur
    return names
cur, agg, tot = None, collections.defaultdict(lambda: [0, 0, 0]), [0, 0, 0]
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        if cur not in fn_of: fn_of[cur] = load(cur)
        continue
    if len(r) > 8 and r[0].isdigit():
        try: s, ie, te = int(r[6] or 0), int(r[7] or 0), int(r[8] or 0)
        except ValueError: continue
        names = fn_of.get(cur)
        key = f"{cur}:{names.get(int(r[0]), '?')}" if names else cur
        a = agg[key]; a[0] += s; a[1] += ie; a[2] += te
        tot[0] += s; tot[1] += ie; tot[2] += te
print(f"total samples {tot[0]}, warp inst {tot[1]}, active {tot[2]/max(tot[1],1):.1f}")
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:40]:
    print(f"{100*a[0]/tot[0]:5.1f}% smp {100*a[1]/tot[1]:5.1f}% inst  act {a[2]/max(a[1],1):5.1f}  {k}")
