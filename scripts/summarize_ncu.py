"""Turns the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_ncu.py r01        # reads gpurun_out/r01_*.{csv,ncu-rep}
"""
import collections, csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "smsp__thread_inst_executed_per_inst_executed.ratio"]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def short(name):
    m = re.search(r"(k_[a-z0-9_]+|Device[A-Za-z]+Kernel)", name)
    return m.group(1) if m else name[:40]


def launches():
    path = os.path.join(G, f"{tag}_launches.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    per = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
        k = short(r[ik])
        per.setdefault(k, []).append(v)
    tot = sum(sum(v) for v in per.values())
    # one-off index build (stl_upload_pack) vs the kernels of the timed step
    build = ("k_kd_refine", "k_leaf_adj", "k_leaf_aabb", "k_inner_aabb", "k_scatter", "k_morton", "k_bbox", "k_index_", "DeviceRadixSort", "at::")
    is_build = lambda k: any(k.startswith(b) or b in k for b in build)
    step_tot = sum(sum(v) for k, v in per.items() if not is_build(k)) or 1.0
    with open(os.path.join(P, f"{tag}_launches.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  (python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras)\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write("# share = of everything the process launched; step% = of the kernels of the timed step (index build excluded)\n")
        f.write(f"{'kernel':34s} {'launches':>8s} {'avg_us':>10s} {'total_us':>11s} {'share':>7s} {'step%':>7s}\n")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            sp = "      -" if is_build(k) else f"{100*sum(v)/step_tot:6.1f}%"
            f.write(f"{k:34s} {len(v):8d} {sum(v)/len(v):10.1f} {sum(v):11.1f} {100*sum(v)/tot:6.1f}% {sp}\n")
    print("wrote", f"{tag}_launches.txt")


def full(name, out, nq_key=None):
    rep = os.path.join(G, f"{tag}_{name}.ncu-rep")
    if not os.path.exists(rep):
        return None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, unit, val = rows[0], rows[1], rows[-1]
    d = {h: (val[i], unit[i]) for i, h in enumerate(hdr)}
    with open(os.path.join(P, out), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on  ({os.path.basename(rep)}, last captured launch)\n")
        f.write(f"kernel: {d.get('Kernel Name', ('?',))[0][:120]}\n")
        for k in KEYS:
            if k in d:
                f.write(f"{k:70s} {d[k][0]:>18s} {d[k][1]}\n")
        st = sorted(((float(v[0].replace(',', '')), h[len(STALLS):]) for h, v in d.items() if h.startswith(STALLS) and "not_issued" not in h and v[0]), reverse=True)
        tot = sum(s for s, _ in st) or 1
        f.write("warp stall samples (all): " + ", ".join(f"{n} {100*s/tot:.1f}%" for s, n in st[:8]) + "\n")
    print("wrote", out)
    return d


launches()
d1 = full("k1_full", f"{tag}_k1_assoc2d.txt")
if d1:
    def num(x):
        v, u = x
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    tr = num(d1["dram__bytes_read.sum"]) + num(d1["dram__bytes_write.sum"])
    json.dump({"kernel": "k_assoc2d", "dram_bytes_per_launch": tr, "source": f"profiles/{tag}_k1_assoc2d.txt (ncu --set full, one launch at the bench shape)"},
              open(os.path.join(P, "k1_traffic.json"), "w"))
    print("k1 traffic", tr)
full("k2_full", f"{tag}_k2_knn.txt")
full("k2b_full", f"{tag}_k2_plane.txt")
full("lm_full", f"{tag}_lm_plane_b.txt" if tag >= "r02" else f"{tag}_lm_knn.txt")
full("lin_full", f"{tag}_linearize.txt")
full("k0_full", f"{tag}_k0_kd_refine.txt")
full("kidx_full", f"{tag}_k0_index_knn.txt")
full("k1poll_full", f"{tag}_k1_poll256.txt")
