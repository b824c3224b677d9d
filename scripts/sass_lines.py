"""Attributes the per-instruction counters of an `ncu --page source --csv` export (SASS rows) to CUDA source lines,
using the line table of the cubin (`nvdisasm --print-line-info`).  python scripts/sass_lines.py CUBIN KERNEL_SUBSTR SRC_CSV [N]"""
import collections, csv, re, subprocess, sys
cubin, kern, src_csv = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout
lines, cur, infn = [], None, False
for ln in out.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        infn = kern in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
rows = list(csv.reader(open(src_csv)))
h = rows[1]
ci, si = h.index("Instructions Executed"), h.index("# Samples")
body = rows[2:]
print(f"{len(lines)} SASS instructions in the cubin, {len(body)} rows in the profile")
agg, samp = collections.Counter(), collections.Counter()
tot = ts = 0
for i, r in enumerate(body):
    key = lines[i] if i < len(lines) else None
    v, s = float(r[ci] or 0), float(r[si] or 0)
    agg[key] += v; samp[key] += s; tot += v; ts += s
srcs = {}
for (k, v) in agg.most_common(topn):
    text = ""
    if k:
        try:
            if k[0] not in srcs:
                import glob
                srcs[k[0]] = open(glob.glob("spatial-temporal-lidar-camera-calibration_b200/csrc/" + k[0])[0]).read().splitlines()
            text = srcs[k[0]][k[1] - 1].strip()[:110]
        except Exception:
            pass
    print(f"  {100 * v / tot:5.1f}% inst {100 * samp[k] / max(ts, 1):5.1f}% samples  {k}  {text}")
