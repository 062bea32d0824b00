#!/usr/bin/env python
"""attribute ncu per-instruction counters to source lines: ncu_lines.py <report.ncu-rep> <kernel mangled-name substring> [launch index]
(joins `ncu --page source --csv` addresses with `nvdisasm --print-line-info` of the cubin inside libssq.so)"""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kname = sys.argv[1], sys.argv[2]
launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
tmp = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s/speedseq_b200/libssq.so >/dev/null 2>&1" % (tmp, ROOT), shell=True)
sass = ""
for f in os.listdir(tmp):
    if f.endswith(".cubin") and kname in subprocess.run("cuobjdump -elf %s/%s" % (tmp, f), shell=True, capture_output=True, text=True).stdout:
        sass = subprocess.run("nvdisasm --print-line-info %s/%s" % (tmp, f), shell=True, capture_output=True, text=True).stdout
lines = sass.split("\n")
start = next(i for i, l in enumerate(lines) if re.match(r"\s*\.section\s+\.text\..*" + re.escape(kname), l) or (l.startswith(".text.") and kname in l))
cur = inl = None
addr2line = {}
for l in lines[start + 1:]:
    if re.match(r"\s*\.section\s+\.text\.", l) and addr2line: break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        chain = re.findall(r'inlined at "([^"]+)", line (\d+)', l)
        inl = [(a.split("/")[-1], int(b)) for a, b in chain]
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: addr2line[int(m.group(1), 16)] = (cur, inl)
out = subprocess.run("ncu -i %s --page source --csv" % rep, shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
blocks.append(len(rows))
b0 = blocks[launch]
h = rows[b0 + 1]; ai = h.index("Address"); ii = h.index("Instructions Executed"); ti = h.index("Thread Instructions Executed"); si = h.index("# Samples")
agg = collections.Counter(); thr = collections.Counter(); smp = collections.Counter(); tot = 0; base = None
mode = os.environ.get("NCU_LINES_MODE", "leaf")  # leaf: innermost source line; top: line in the kernel's own file
for r in rows[b0 + 2:blocks[launch + 1]]:
    if len(r) <= ii or not r[ai]: continue
    a = int(r[ai], 16) if r[ai].startswith("0x") else int(r[ai])
    if base is None: base = a
    cur, inl = addr2line.get(a - base, (("?", 0), None))
    key = cur
    if mode == "top" and inl: key = inl[-1]
    n = int(r[ii] or 0); agg[key] += n; thr[key] += int(r[ti] or 0); smp[key] += int(r[si] or 0); tot += n
print("kernel", rows[b0][1][:60], "warp-instructions", tot)
srcs = {}
def src(k):
    if k[0] not in srcs:
        p = os.path.join(ROOT, "speedseq_b200/csrc", k[0])
        srcs[k[0]] = open(p).read().split("\n") if os.path.exists(p) else []
    s = srcs[k[0]]
    return s[k[1] - 1].strip()[:120] if 0 < k[1] <= len(s) else ""
for k, v in agg.most_common(int(os.environ.get("NCU_LINES_TOP", "45"))):
    print("%5.1f%% inst  %4.1f lanes  %6d smp  %s:%d  %s" % (100.0 * v / tot, thr[k] / max(1, v), smp[k], k[0], k[1], src(k)))
