#!/usr/bin/env python
"""Attribute an ncu report's per-instruction samples to CUDA source lines.

usage: ncu_lines.py report.ncu-rep kernel_substring libfile.so [top_n]
Uses `ncu --page source --csv` (SASS rows in address order) and `nvdisasm --print-line-info` of the
cubin embedded in the library (same order), so it works for -lineinfo builds without the GUI."""
import csv, collections, os, re, subprocess, sys, tempfile

rep, kern, lib = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
lines = []   # (line number) per SASS instruction of the kernel, in order
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    out = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur, inside, lst = None, False, []
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            if inside and lst:
                break
            inside = kern in m.group(1)
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            if "/csrc/" in m.group(1) or m.group(1).startswith("qiw_"):
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            lst.append(cur)
    if lst:
        lines = lst
        break
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]
c_s, c_i, c_w = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared")
sass = rows[hi + 1:]
print("SASS rows in report %d, instructions with line info %d" % (len(sass), len(lines)))
agg = collections.defaultdict(lambda: [0, 0, 0])
for k, r in enumerate(sass):
    ln = lines[k] if k < len(lines) else None
    a = agg[ln]
    a[0] += int(r[c_s] or 0); a[1] += int(r[c_i] or 0); a[2] += int(r[c_w] or 0)
ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
csrc = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc")
srcs = {}
def text_of(ln):
    if not ln:
        return "?"
    f, n = ln
    if f not in srcs:
        try:
            srcs[f] = open(os.path.join(csrc, f)).read().splitlines()
        except OSError:
            srcs[f] = []
    return srcs[f][n - 1].strip()[:110] if 0 < n <= len(srcs[f]) else "?"
print("samples %d  warp instructions %d" % (ts, ti))
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    where = "%s:%d" % (ln[0].replace("qiw_", "").replace(".cuh", "").replace(".cu", ""), ln[1]) if ln else "?"
    print("%5.1f%% samples %5.1f%% inst  smem wavefronts %11d | %12s: %s" % (100 * a[0] / ts, 100 * a[1] / ti, a[2], where, text_of(ln)))
