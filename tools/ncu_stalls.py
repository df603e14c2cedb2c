#!/usr/bin/env python
"""Warp-stall samples of one kernel by source line and by device function.

    ncu -i prof.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all libbayeslands_b200.so ; nvdisasm --print-line-info bl_v1024.sm_100a.cubin > v1024.dis
    python tools/ncu_stalls.py src.csv v1024.dis <kernel-name-fragment> [top]

Joins the SASS rows of the ncu source page (offset = address - first address) with nvdisasm's `//## File .. line ..`
annotations of the same build (compile with -lineinfo).
"""
import collections
import csv
import re
import sys

src_csv, dis, frag = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45

# ---- nvdisasm: offset -> (file, line) within the kernel's .text section
line_of = {}
cur = None
inside = False
for ln in open(dis, errors="replace"):
    if ln.startswith(".text."):
        inside = frag in ln
        cur = None
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())

rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = None
by_line = collections.defaultdict(lambda: collections.Counter())
total = collections.Counter()
inst_line = collections.Counter()
local = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    addr = int(r[col["Address"]], 16)
    if base is None:
        base = addr
    off = addr - base
    key, sass = line_of.get(off, (None, ""))
    key = key or ("?", 0)
    n = int(r[col["# Samples"]] or 0)
    ie = int(r[col["Instructions Executed"]] or 0)
    inst_line[key] += ie
    if r[col["Address Space"]].strip().lower() == "local":
        local[key] += ie
    for s in stall_cols:
        v = int(r[col[s]] or 0)
        if v:
            by_line[key][s] += v
            total[s] += v
tot = sum(total.values())
print("samples %d; by reason: %s" % (tot, ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in total.most_common(9))))
print("warp instructions executed %.1f M, of them local-memory (spill) accesses %.1f M" % (sum(inst_line.values()) / 1e6, sum(local.values()) / 1e6))
srcs = {}
def text(f, l):
    if f not in srcs:
        try:
            import os
            here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
            srcs[f] = open(os.path.join(here, "bayeslands_b200", "csrc", f)).read().split("\n")
        except Exception:
            srcs[f] = []
    s = srcs[f]
    return s[l - 1].strip()[:70] if 0 < l <= len(s) else ""
print("\nfile:line                 samples  warp inst  local   top reasons                    source")
for key, cnt in sorted(by_line.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    n = sum(cnt.values())
    rs = ", ".join("%s %d%%" % (k[6:], round(100.0 * v / n)) for k, v in cnt.most_common(2))
    print("%-24s %6.2f%% %9.1fM %6.1fM  %-30s %s" % ("%s:%d" % key, 100.0 * n / tot, inst_line[key] / 1e6, local[key] / 1e6, rs, text(*key)))
# ---- by device function: ranges of source lines (function starts are found with a crude scan of the sources)
print("\nspill (local-memory) warp instructions by source line, top 25:")
for key, v in local.most_common(25):
    print("  %-24s %8.1fM   %s" % ("%s:%d" % key, v / 1e6, text(*key)))
