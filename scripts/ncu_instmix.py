"""Dynamic instruction mix of a kernel from an ncu report:  python scripts/ncu_instmix.py report.ncu-rep [sass.csv]

Reads `ncu --page source --csv --print-source sass` (or a saved dump), groups the SASS opcodes into classes and
sums, per class, warp-level instructions executed, predicated-on thread instructions and PC samples.  Answers
"what share of the issue slots is arithmetic the algorithm asks for, and what is bookkeeping"."""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

CLASSES = [
    ("fp32 fma/mul/add", r"^(FFMA|FMUL|FADD|FFMA32I|FMUL32I|FADD32I)\b"),
    ("fp32 min/max/compare/select", r"^(FMNMX|FSETP|FSEL|FSET|FCHK)\b"),
    ("special function (MUFU)", r"^MUFU\b"),
    ("conversion", r"^(F2I|I2F|F2F|I2FP|F2FP|FRND|I2I)\b"),
    ("integer / logic / shift", r"^(IADD3|IADD|IMAD|IMUL|LOP3|LOP|SHF|SHL|SHR|LEA|IABS|IMNMX|POPC|FLO|BREV|PRMT|SGXT|BMSK|VIADD|VIMNMX|IDP|VABSDIFF)\b"),
    ("integer compare / predicate / select", r"^(ISETP|PLOP3|SEL|P2R|R2P|PSETP|CSET|ICMP)\b"),
    ("move", r"^(MOV|MOV32I|UMOV|S2R|S2UR|CS2R|R2UR|UR2R)\b"),
    ("uniform datapath", r"^U[A-Z0-9]+\b"),
    ("shared memory", r"^(LDS|STS|LDSM|ATOMS)\b"),
    ("global / local / constant memory", r"^(LDG|STG|LD|ST|LDL|STL|LDC|LDCU|ATOM|ATOMG|RED|CCTL|MEMBAR|ERRBAR)\b"),
    ("warp vote / shuffle / match", r"^(VOTE|VOTEU|SHFL|MATCH|REDUX|WARPSYNC|NANOSLEEP)\b"),
    ("barrier", r"^(BAR|DEPBAR|B2R|R2B)\b"),
    ("branch / call / convergence", r"^(BRA|BRX|JMP|JMX|CALL|RET|EXIT|BSSY|BSYNC|BREAK|BMOV|YIELD|KILL|NOP|BPT)\b"),
]


def opcode(src):
    s = src.strip()
    s = re.sub(r"^@!?U?P\w+\s+", "", s)  # predicate guard
    return s.split()[0].rstrip(";") if s else ""


def main():
    if len(sys.argv) > 2:
        text = open(sys.argv[2]).read()
    else:
        text = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                              capture_output=True, text=True, check=True).stdout
    lines = text.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    print(lines[start - 1].split('","')[1].rstrip('",') if start else "")
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    acc = defaultdict(lambda: [0, 0, 0, 0, defaultdict(int)])  # static, warp inst, thread inst (pred on), samples, per-op
    for r in rows:
        op = opcode(r["Source"])
        base = op.split(".")[0]
        cls = next((name for name, pat in CLASSES if re.match(pat, base)), "other")
        a = acc[cls]
        a[0] += 1
        a[1] += int(r["Instructions Executed"] or 0)
        a[2] += int(r["Predicated-On Thread Instructions Executed"] or 0)
        a[3] += int(r["# Samples"] or 0)
        a[4][base] += int(r["Instructions Executed"] or 0)
    tw = sum(a[1] for a in acc.values())
    tt = sum(a[2] for a in acc.values())
    ts = sum(a[3] for a in acc.values())
    print(f"{len(rows)} SASS instructions, {tw:,} warp instructions executed, {tt:,} thread instructions "
          f"({tt / tw:.1f} of 32 lanes on average), {ts:,} PC samples\n")
    print("| class | static | warp instructions | share | lanes on | PC samples | top opcodes |")
    print("|---|---:|---:|---:|---:|---:|---|")
    for cls, a in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        top = ", ".join(f"{o} {100 * n / tw:.1f}%" for o, n in sorted(a[4].items(), key=lambda kv: -kv[1])[:4] if n)
        print(f"| {cls} | {a[0]} | {a[1]:,} | {100 * a[1] / tw:.1f}% | {a[2] / max(a[1], 1):.1f} | {100 * a[3] / max(ts, 1):.1f}% | {top} |")


if __name__ == "__main__":
    main()
