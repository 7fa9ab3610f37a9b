#!/usr/bin/env python
"""profiles/ptxas_resources.md from the `-Xptxas -v` logs the Makefile leaves next to the objects
(vecchio_b200/csrc/ptxas_*.log).  Run after `make`:  python profiles/make_ptxas_resources.py"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for log in sorted(glob.glob(os.path.join(ROOT, "vecchio_b200", "csrc", "ptxas_*.log"))):
    build = os.path.basename(log)[len("ptxas_"):-len(".log")]
    text = open(log).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                         r"ptxas info\s*: Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", text):
        sym, stack, st, ld, regs, _, smem = m.groups()
        name = subprocess.run(["c++filt", sym], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)  # drop the parameter list
        rows.append((build, name, regs, stack, f"{st}/{ld}", smem or "0"))
out = ["# ptxas resource usage of every kernel (`nvcc -Xptxas -v`, sm_100a), from `vecchio_b200/csrc/ptxas_*.log`", "",
       "Made by `profiles/make_ptxas_resources.py` after `make`.  Builds: `fast` / `l0` / `strict` = `vk_kernels.cu` as `vkfast`, `vkfast_l0`",
       "(one unflipped Rect light, no SpecDiffuse) and `vkstrict`; `warpq_*`, `stepq_*`, `staged_*`, `wf_*` likewise; `*_simple` = the trimmed",
       "build `vkfast_simple` for scenes that reach no non-solid texture, Metal, SpecDiffuse, second light, moving sphere or (u, v): Cornell box,",
       "Cornell smoke.  Shared memory of the queue kernels is dynamic (not shown by ptxas).", "",
       "| build | kernel | registers | stack B | spill st/ld B | smem B |", "|---|---|---|---|---|---|"]
out += [f"| {b} | `{n}` | {r} | {s} | {sp} | {sm} |" for b, n, r, s, sp, sm in rows]
open(os.path.join(ROOT, "profiles", "ptxas_resources.md"), "w").write("\n".join(out) + "\n")
print(f"{len(rows)} kernels")
