#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brings back into the tracked summaries under profiles/.

    python profiles/summarize.py full   <report.ncu-rep> <out.md>      # one `ncu --set full` capture
    python profiles/summarize.py launch <launches.csv>   <out.md>      # the gpu__time_duration launch list

Runs in the CPU container (ncu -i reads reports without a GPU).  Nothing here is on the product path.
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

RAW_KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), CTAs/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (% of 64 warps/SM)"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "warp execution efficiency: active threads per warp instruction (of 32)"),
    ("smsp__thread_inst_executed_pred_on_per_inst_executed.ratio", "  ... predicated-on threads per warp instruction"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots used (% of peak)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FP32 FMA pipe issue fraction (% of peak)"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active (%)"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe (%)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (transcendental / conversion) pipe (%)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe (%)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput (% of peak)"),
    ("l1tex__t_sector_hit_rate.pct", "L1/TEX hit rate (%)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate (%)"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput (% of peak)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput (% of peak)"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)"),
    ("dram__bytes_read.sum", "DRAM bytes read"), ("dram__bytes_write.sum", "DRAM bytes written"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__thread_inst_executed.sum", "thread instructions executed"),
    ("sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "FFMA thread instructions"),
    ("sm__sass_thread_inst_executed_op_fadd_pred_on.sum", "FADD thread instructions"),
    ("sm__sass_thread_inst_executed_op_fmul_pred_on.sum", "FMUL thread instructions"),
    ("smsp__inst_executed_op_local_ld.sum", "local-memory (spill/stack) loads, warp instructions"),
    ("smsp__inst_executed_op_local_st.sum", "local-memory (spill/stack) stores, warp instructions"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "branch_resolving", "no_instruction", "not_selected",
          "dispatch_stall", "lg_throttle", "mio_throttle", "tex_throttle", "barrier", "membar", "sleeping", "drain", "misc", "selected"]


def ncu_csv(report, page, extra=()):
    out = subprocess.run(["ncu", "-i", report, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def full(report, out_path):
    rows = ncu_csv(report, "raw")
    hdr, units, launches = rows[0], rows[1], rows[2:]
    lines = [f"# ncu --set full summary of `{os.path.basename(report)}`", "",
             "Made by `profiles/summarize.py full` from the report a `gpurun` call brought back (command in the header of the",
             "round's job script under `scripts/`).  Cold-cache, serialised replays: use the ratios, not the absolute time.", ""]
    summary = []
    for v in launches:
        name = v[hdr.index("Kernel Name")]
        lines += [f"## `{name.split('(')[0]}`", "", "| metric | value |", "|---|---|"]
        js = {"kernel": name.split("(")[0]}
        for key, label in RAW_KEYS:
            if key in hdr:
                i = hdr.index(key)
                lines.append(f"| {label} (`{key}`) | {v[i]} {units[i]} |")
                js[key] = v[i] + " " + units[i]
        try:
            rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            js["dram_bytes_per_launch"] = float(v[rd].replace(",", "")) * mult[units[rd]] + float(v[wr].replace(",", "")) * mult[units[wr]]
        except Exception:
            pass
        lines += ["", "Warp stall reasons (warps stalled per issue-active cycle, `smsp__average_warps_issue_stalled_*_per_issue_active.ratio`):", ""]
        st = []
        for s in STALLS:
            k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if k in hdr:
                try:
                    st.append((float(v[hdr.index(k)]), s))
                except ValueError:
                    pass
        lines.append(", ".join(f"{s} {x:.2f}" for x, s in sorted(st, reverse=True) if x >= 0.005))
        lines.append("")
        summary.append(js)
    # source page: aggregate by source line
    src = ncu_csv(report, "source", ("--print-source", "cuda,sass"))
    cur, agg, tot = None, collections.defaultdict(lambda: [0, 0, 0, ""]), [0, 0, 0]
    for r in src:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) > 8 and r[0].isdigit():
            try:
                s, ie, te = int(r[6] or 0), int(r[7] or 0), int(r[8] or 0)
            except ValueError:
                continue
            a = agg[(cur, int(r[0]))]
            a[0] += s; a[1] += ie; a[2] += te; a[3] = r[1].strip()[:100]
            tot[0] += s; tot[1] += ie; tot[2] += te
    if tot[1]:
        lines += ["## Hottest source lines (PC samples; warp instructions; active threads per instruction)", "",
                  f"total: {tot[0]} samples, {tot[1]} warp instructions, {tot[2] / tot[1]:.2f} active threads per warp instruction", "",
                  "| samples | warp inst | active | line | source |", "|---|---|---|---|---|"]
        for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:40]:
            code = a[3].replace("|", "\\|")
            lines.append(f"| {100 * a[0] / tot[0]:.1f}% | {100 * a[1] / tot[1]:.1f}% | {a[2] / max(a[1], 1):.1f} | {k[0]}:{k[1]} | `{code}` |")
        by_file = collections.defaultdict(lambda: [0, 0, 0])
        for k, a in agg.items():
            b = by_file[k[0]]
            b[0] += a[0]; b[1] += a[1]; b[2] += a[2]
        lines += ["", "| file | samples | warp inst | active |", "|---|---|---|---|"]
        for f, b in sorted(by_file.items(), key=lambda x: -x[1][0])[:8]:
            lines.append(f"| {f} | {100 * b[0] / tot[0]:.1f}% | {100 * b[1] / tot[1]:.1f}% | {b[2] / max(b[1], 1):.1f} |")
    open(out_path, "w").write("\n".join(lines) + "\n")
    json.dump(summary, open(os.path.splitext(out_path)[0] + ".json", "w"), indent=1)
    print("wrote", out_path)


def launch(path, out_path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0, r[hdr.index("Grid Size")], r[hdr.index("Block Size")]])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    lines = [f"# ncu launch list (`gpu__time_duration.sum`, `--clock-control none`) from `{os.path.basename(path)}`", "",
             "Per-launch times are cold-cache and serialised: the SHARE of the step is what must agree with bench.py.", "",
             "| kernel | launches | total ms | share | avg ms | grid | block |", "|---|---|---|---|---|---|---|"]
    for k, a in agg.items():
        lines.append(f"| `{k}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot:.1f}% | {a[1] / a[0] / 1e6:.4f} | {a[2]} | {a[3]} |")
    open(out_path, "w").write("\n".join(lines) + "\n")
    print("wrote", out_path)


if __name__ == "__main__":
    {"full": full, "launch": launch}[sys.argv[1]](sys.argv[2], sys.argv[3])
