"""The SASS of the (first) kernel in an ncu report grouped into basic blocks -- runs of instructions with the same
execution count and the same active lanes -- with executions normalised per 32 units of work, ranked by cost.
usage: python scripts/ncu_sass_blocks.py report.ncu-rep units [top]     (units: e.g. ray segments traced by the launch)
Reads the report with `ncu -i ... --page source --print-source sass --csv` (works without a GPU)."""
import csv, io, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
name = rows[0][1] if rows and len(rows[0]) > 1 else "?"
rows = [r for r in rows[2:] if len(r) > 9 and r[0].startswith("0x")]
per = units / 32.0
blocks, cur = [], None
for i, r in enumerate(rows):
    c, smp, act = int(r[5]) / per, int(r[4]), float(r[8] or 0)
    if cur and abs(cur["c"] - c) < 1e-9 and abs(cur["a"] - act) < 0.01:
        cur["n"] += 1; cur["s"] += smp; cur["end"] = i
    else:
        cur = {"start": i, "end": i, "c": c, "a": act, "n": 1, "s": smp, "first": r[1].strip()}
        blocks.append(cur)
tot = sum(b["c"] * b["n"] for b in blocks); ts = sum(b["s"] for b in blocks) or 1
print(f"# SASS basic blocks of `{name.split('(')[0]}`\n")
print(f"{len(rows)} SASS instructions; {tot:.1f} warp instructions executed per 32 units of work ({units:.0f} units in the launch); {ts} PC samples.\n")
print("| instructions | executions per 32 units | static size | warp instructions per 32 units | share | PC samples | lanes on | first instruction |")
print("|---|---:|---:|---:|---:|---:|---:|---|")
for b in sorted(blocks, key=lambda b: -b["c"] * b["n"])[:top]:
    print(f"| {b['start']}-{b['end']} | {b['c']:.3f} | {b['n']} | {b['c'] * b['n']:.1f} | {100 * b['c'] * b['n'] / tot:.1f} % | {100 * b['s'] / ts:.1f} % | {b['a']:.1f} | `{b['first'][:60]}` |")
