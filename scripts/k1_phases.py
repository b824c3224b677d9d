"""Buckets the per-instruction counters of an `ncu --page source --csv --print-source sass` export of k_assoc2d by PHASE of the
kernel (outermost source line of every SASS instruction, from the cubin's line table incl. inlining).
python scripts/k1_phases.py CUBIN SRC_CSV"""
import collections, csv, re, subprocess, sys
cubin, src_csv = sys.argv[1:3]
kern = sys.argv[3] if len(sys.argv) > 3 else "k_assoc2d"
src_file = sys.argv[4] if len(sys.argv) > 4 else "assoc2d.cu"
out = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout
lines, cur, infn = [], None, False
for ln in out.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        infn = kern in m.group(1)
        continue
    if not infn:
        continue
    if "//## File" in ln:
        # outermost = last (file, line) pair on the row that lies in the kernel's own file
        pairs = re.findall(r'"([^"]+)", line (\d+)', ln)
        own = [(f, int(l)) for f, l in pairs if f.endswith(src_file)]
        cur = own[-1][1] if own else None
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
# line ranges of assoc2d.cu (update when the file moves): device helpers are attributed to the phase that inlines them
PH = [("prologue (thread 0)", 197, 302), ("unit setup", 303, 326), ("cells (A1) + table wait", 327, 366), ("cells (A1) + table wait", 125, 140),
      ("stream (A2)", 367, 451), ("exact-1", 452, 470), ("exact-1", 87, 100), ("exact-1", 156, 193), ("tie pass", 471, 504),
      ("corrset", 505, 555), ("covisible pairs", 556, 588), ("reduction + epilogue", 589, 618), ("hand-eye", 104, 120)]
def phase(l):
    if l is None: return "?"
    for n, a, b in PH:
        if a <= l <= b: return n
    return f"line {l}"
rows = list(csv.reader(open(src_csv)))
h = rows[1]
ci, si = h.index("Instructions Executed"), h.index("# Samples")
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
inst, samp, st = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
for i, r in enumerate(rows[2:]):
    p = phase(lines[i] if i < len(lines) else None)
    inst[p] += float(r[ci] or 0); samp[p] += float(r[si] or 0)
    for j, c in stall_cols:
        st[p][c] += float(r[j] or 0)
ti, ts = sum(inst.values()), sum(samp.values())
print(f"{'phase':14s} {'inst%':>6s} {'samples%':>8s}  top stalls")
for p, _ in samp.most_common():
    tops = ", ".join(f"{c[6:]} {100*v/max(samp[p],1):.0f}%" for c, v in st[p].most_common(4))
    print(f"{p:14s} {100*inst[p]/ti:6.1f} {100*samp[p]/ts:8.1f}  {tops}")
