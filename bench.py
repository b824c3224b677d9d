#!/usr/bin/env python
"""bench.py — extrinsic cost evaluations / s at KITTI-00 shape (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One STEP = one complete evaluation of ONE candidate extrinsic over the whole keyframe set
(BASELINE.json configs[1]: ~1500 keyframes x ~117k-point scans x 2000 keypoints):
    BAError cost           stl_eval_batch   (iba_global.cpp:169-344)
  + association at x       stl_associate    (BuildProblem, iba_local.cpp:145-323)
  + cost / J^T r / 7x7 J^T J  stl_linearize_batch (IBACalib2.hpp factors + Huber)
`value`  : steps/s with the pack resident in HBM, results left on the device, timed with CUDA
           events on the launching stream (max over ranks).
`e2e`    : the same step through the host-facing C-ABI calls (host x in, host results out,
           H2D/D2H and syncs inside the timed region).
N > 1    : one process per GPU.  Default sharding is by CANDIDATE (every rank holds the whole
           pack and evaluates its own candidate each step — independent units, no data-path
           collective => weak scaling).  `--shard keyframes` runs the north-star layout instead:
           keyframes sharded, one NCCL fp64 all-reduce of the [1,12]+[1,61] records per step.
The CPU baseline (`cpu_baseline`, and the whole `--impl reference` arm) is the oracle — the
reference's vendored nanoflann (oracle/_ref) when it was built, else the in-repo port — timed on
the host cores with OpenMP over keyframes (iba_func.cpp:203,463) on a bounded keyframe sample.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "spatial-temporal-lidar-camera-calibration_b200"
METRIC = "extrinsic_cost_evals_per_s"
UNIT = "evals/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nkf", type=int, default=1500, help="keyframes of the sequence (KITTI-00 shape: 1500)")
    ap.add_argument("--shard", default="candidates", choices=["candidates", "keyframes"])
    ap.add_argument("--cpu-sample-kf", type=int, default=192, help="keyframes of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-poll-batch", action="store_true", help="skip the 256-candidate poll-batch measurement")
    ap.add_argument("--poll-batch", type=int, default=256)
    ap.add_argument("--no-plane-index", action="store_true", help="skip the extra measurement with params.plane_index=1")
    ap.add_argument("--seed", type=int, default=1000)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML is polled
    from a thread every 2 ms (the timed region is ~0.1 s, too short for `nvidia-smi -lms`); nvidia-smi
    is the fallback when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.h, self.stop_flag, self.sm, self.mask = None, None, False, [], 0
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                self.mask |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1.0)
            try:
                mx = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.h, self.nvml.NVML_CLOCK_SM))
            except Exception:
                mx = None
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": mx,
                    "reasons": sorted(v for b, v in self.BITS.items() if self.mask & b), "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def oracle_sample_rate(pack, X, nthreads=0, min_seconds=8.0, want_lm=True):
    """Times the CPU oracle on `pack` (a keyframe sample): seconds per (candidate x sample) step."""
    from oracle import oracle as O
    kind = "ref" if O.have_ref() else "port"
    if nthreads <= 0:
        nthreads = os.cpu_count() or 1
    orc = O.Oracle(pack, kind=kind, nthreads=nthreads)
    cores = nthreads
    reps, t_tot, i = 0, 0.0, 0
    orc.ba_error_sums(X[:1], mode=1, strict=True, nthreads=nthreads)  # warm-up
    while t_tot < min_seconds and reps < 200:
        x = X[i % len(X)]
        t0 = time.perf_counter()
        orc.ba_error_sums(x, mode=1, strict=True, nthreads=nthreads)  # BAError, OpenMP over keyframes (iba_func.cpp:203)
        if want_lm:
            orc.associate(x, strict=True)                   # BuildProblem (OpenMP over keyframes, iba_local.cpp:162)
            orc.linearize(x)                                # one evaluation of all residual blocks + Jacobians
        t_tot += time.perf_counter() - t0
        reps += 1
        i += 1
    return t_tot / reps, cores, kind, orc.build_seconds, reps


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle; real nanoflann when built) on the host cores."""
    if rank != 0:
        return
    synth = importlib.import_module(PKG + ".synth")
    nsamp = min(args.cpu_sample_kf, args.nkf)
    pack, x_gt, _ = synth.generate(n_kf=nsamp, n_kf_total=args.nkf, seed=args.seed)
    X = synth.candidates(x_gt, max(args.steps + args.warmup, 2), 0.2)
    from oracle import oracle as O
    kind = "ref" if O.have_ref() else "port"
    cores = os.cpu_count() or 1
    orc = O.Oracle(pack, kind=kind, nthreads=cores)

    def step(x):
        orc.ba_error_sums(x, mode=1, strict=True, nthreads=cores)
        orc.associate(x, strict=True)
        orc.linearize(x)
    for i in range(args.warmup):
        step(X[i])
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(X[args.warmup + i])
    dt = (time.perf_counter() - t0) / args.steps
    scale = nsamp / args.nkf
    val = scale / dt
    sample = (f"{nsamp} of {args.nkf} keyframes per step (extrapolated linearly to {args.nkf}); OpenMP over keyframes, "
              f"{cores} threads; KNN = {'reference nanoflann v1.5.0 (oracle/_ref)' if kind == 'ref' else 'in-repo nanoflann port'}; "
              f"one-off KD-tree build {orc.build_seconds:.2f}s excluded")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3 / scale, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"KITTI-00 shape: {args.nkf} keyframes x ~117k-pt scans x 2000 keypoints, 1 candidate/step: BAError cost + association + cost/JtJ"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference" if kind == "ref" else "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncpu = os.cpu_count() or 8
    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1; the CPU arm is entitled to every host core (rank 0 runs alone)
        os.environ["OMP_NUM_THREADS"] = str(ncpu)
        run_reference(args, rank, world)
        return
    os.environ["OMP_NUM_THREADS"] = str(max(1, ncpu // max(world, 1)))  # host-side generator / pack preparation only

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the cost-evaluation path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    synth = importlib.import_module(PKG + ".synth")
    capi = importlib.import_module(PKG + ".capi")
    par = importlib.import_module(PKG + ".parallel")
    _abi = importlib.import_module(PKG + "._abi")

    F = args.nkf
    by_kf = world > 1 and args.shard == "keyframes"
    t0 = time.time()
    if by_kf:
        kb, ke = par.shard_bounds(F, world, rank)
        pack, x_gt, _ = synth.generate(n_kf=ke - kb, kf_begin=kb, n_kf_total=F, seed=args.seed)
    else:
        pack, x_gt, _ = synth.generate(n_kf=F, seed=args.seed)
    t_gen = time.time() - t0
    nsteps = args.steps + args.warmup
    # every (rank, step) gets its own candidate; row 0 of the list is the ground truth
    Xall = synth.candidates(x_gt, nsteps * world + 1, 0.2, seed=42)
    X = Xall[1:][rank::world] if not by_kf else Xall[1: nsteps + 1]

    ctx = capi.Context(device=local_rank)
    t0 = time.time()
    ctx.upload(pack)
    torch.cuda.synchronize()
    t_upload = time.time() - t0
    build_ms = ctx.stage_stats()["build"][0]
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    d_sums = torch.zeros((1, _abi.STL_EVAL_NSUMS), dtype=torch.float64, device=dev)
    d_lin = torch.zeros((1, _abi.STL_LIN_NSUMS), dtype=torch.float64, device=dev)

    def step_device(x):
        ctx.eval_sums_device(x, d_sums.data_ptr(), stream.cuda_stream)
        ctx.associate(x)
        ctx.linearize_device(x, d_lin.data_ptr(), stream.cuda_stream)
        if by_kf:  # the only exchange of the path: per-candidate cost record + normal equations
            dist.all_reduce(d_sums)
            dist.all_reduce(d_lin)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step_device(X[i])
    barrier()
    launches0 = ctx.work_counters()["launches"]
    ctx.set_profiling(True)
    ctx.stage_stats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for i in range(args.steps):
        step_device(X[args.warmup + i])
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    stats = ctx.stage_stats()
    ctx.set_profiling(False)
    launches = int(ctx.work_counters()["launches"] - launches0)
    wc = ctx.work_counters()
    sums_last = d_sums.cpu().numpy()[0]
    lin_last = d_lin.cpu().numpy()[0]

    # ---- end to end through the host-facing C-ABI (host buffers in and out)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        x = X[args.warmup + i]
        s = ctx.eval_sums(x)
        ctx.associate(x)
        L = ctx.linearize(x)
        if by_kf:
            s = par.SumAllReduce(dev)(s)
            L = par.SumAllReduce(dev)(L)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    # ---- BASELINE configs[3] shape: a NOMAD poll batch of 256 candidates (cost) + the same batch linearised on the
    # frozen association (cost + JtJ), candidates dealt out across the ranks; device-timed, reported beside `value`
    pb_ms, pb_B = float("nan"), 0
    if not args.no_poll_batch:
        pb_total = args.poll_batch
        if by_kf:
            Xp = synth.candidates(x_gt, pb_total + 1, 0.2, seed=43)[1:]
        else:
            Xp = synth.candidates(x_gt, pb_total + 1, 0.2, seed=43)[1:][rank::world]
        pb_B = len(Xp)
        d_ps = torch.zeros((pb_B, _abi.STL_EVAL_NSUMS), dtype=torch.float64, device=dev)
        d_pl = torch.zeros((pb_B, _abi.STL_LIN_NSUMS), dtype=torch.float64, device=dev)

        def poll_step():
            ctx.eval_sums_device(Xp, d_ps.data_ptr(), stream.cuda_stream)
            ctx.linearize_device(Xp, d_pl.data_ptr(), stream.cuda_stream)
            if by_kf:
                dist.all_reduce(d_ps)
                dist.all_reduce(d_pl)
        ctx.associate(Xp[0])
        poll_step()                                                          # untimed: sizes the batch buffers (pinned + device)
        barrier()
        qe0, qe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        qe0.record(stream)
        for _ in range(2):
            poll_step()
        qe1.record(stream)
        barrier()
        pb_ms = qe0.elapsed_time(qe1) / 2

    # ---- the same step with the optional plane index (local planes fitted once at upload, looked up after)
    pi_ms, pi_upload = float("nan"), float("nan")
    if not args.no_plane_index:
        ctx.close()
        pkgmod = importlib.import_module(PKG)
        pp = pkgmod.default_params()
        pp.plane_index = 1
        ctx = capi.Context(params=pp, device=local_rank)
        t0 = time.time()
        ctx.upload(pack)
        torch.cuda.synchronize()
        pi_upload = time.time() - t0
        ctx.set_stream(stream.cuda_stream)
        for i in range(args.warmup):
            step_device(X[i])
        barrier()
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record(stream)
        for i in range(args.steps):
            step_device(X[args.warmup + i])
        pe1.record(stream)
        barrier()
        pi_ms = pe0.elapsed_time(pe1)
        pi_check = float(d_sums.cpu().numpy()[0][0])

    tmax = torch.tensor([ms, e2e_s * 1e3, pi_ms, pb_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max, pi_ms_max, pb_ms_max = float(tmax[0]), float(tmax[1]), float(tmax[2]), float(tmax[3])
    units_per_step = 1 if by_kf else world
    value = units_per_step * args.steps / (ms_max * 1e-3)
    e2e_value = units_per_step * args.steps / (e2e_ms_max * 1e-3)

    if rank == 0:
        peak, peak_src = peaks()
        k1_ms, k1_n = stats["assoc2d"]
        k1_avg = k1_ms / max(k1_n, 1)
        n_pts, n_kp = pack.n_points, pack.n_keypoints
        k1_bytes = 12.0 * n_pts + 16.0 * n_kp  # per launch: one candidate over this rank's keyframes
        achieved = k1_bytes / (k1_avg * 1e-3) / 1e9 if k1_avg > 0 else 0.0
        tot_stage = sum(v[0] for k, v in stats.items() if k != "build")
        share = {k: round(v[0] / tot_stage, 4) for k, v in stats.items() if v[1] and k != "build"}
        k2_ms, k2_n = stats["knn3d"]
        q3 = float(sums_last[6]) if not by_kf else float(sums_last[6])
        knn_q_eval = n_kp + 2.0 * q3  # findNeighbors calls of one BAError (iba_global.cpp:92,120,129)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong" if by_kf else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"KITTI-00 shape: {F} keyframes x ~117k-pt 64-beam scans x 2000 keypoints, 1 candidate/step/GPU: "
                            "BAError cost + association + cost/JtJ (BASELINE configs[1])",
                "keyframes": F, "points": int(n_pts), "keypoints": int(n_kp), "candidates_per_step": units_per_step,
                "sharding": ("keyframes + NCCL allreduce" if by_kf else ("candidates (pack replicated per GPU)" if world > 1 else "single GPU")),
                "l2_policy": f"inputs larger than L2 ({12.0 * n_pts / 1e9:.2f} GB of scans streamed per evaluation)",
                "upload_s": round(t_upload, 3), "index_build_ms": round(build_ms, 1), "synth_s": round(t_gen, 2),
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 208 * 2 + 1608, "d2h_bytes_per_step": 96 + 488 + 12},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {
                "kernel": "k_assoc2d (K1: transform + project + 2-D association)", "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": k1_bytes, "avg_launch_ms": k1_avg, "launches": k1_n,
            },
            "stage_share": share,
            "stage_ms_per_launch": {k: round(v[0] / v[1], 4) for k, v in stats.items() if v[1] and k != "build"},
            "knn": {"queries_per_eval": knn_q_eval, "queries_per_s_whole_step": knn_q_eval * value,
                    "k2_pairs_per_s": (q3 / (k2_ms / max(k2_n, 1) * 1e-3)) if k2_ms > 0 else None},
            "poll_batch": None if args.no_poll_batch else {
                "candidates": args.poll_batch, "per_rank": pb_B, "ms": pb_ms_max,
                "evals_per_s": args.poll_batch / (pb_ms_max * 1e-3), "unit": UNIT,
                "note": "BASELINE configs[3]: one stl_eval_batch (BAError sums) + one stl_linearize_batch (cost, J^T r, J^T J on the "
                        "frozen association) over the whole poll batch; device-timed mean of 2 polls after one untimed poll, max over ranks"},
            "plane_index_option": None if args.no_plane_index else {
                "value": units_per_step * args.steps / (pi_ms_max * 1e-3), "unit": UNIT, "ms_per_step": pi_ms_max / args.steps,
                "upload_s": round(pi_upload, 3), "extra_hbm_bytes": 36 * int(n_pts),
                "note": "params.plane_index=1: candidate-independent plane fits moved into the index build; identical results "
                        "(tests/test_gpu_parity.py::test_plane_index_option_gives_identical_results). Reported beside `value`, not as it.",
                "f_sum_check": pi_check},
            "result_check": {"f_sums": [float(sums_last[0]), float(sums_last[1])], "lm_cost": float(lin_last[0]),
                             "frames_kept": float(sums_last[10])},
        }
        # the ncu traffic figure, if a capture of this kernel has been summarised under profiles/
        tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if os.path.exists(tp):
            try:
                line["roofline"]["traffic"] = json.load(open(tp))["dram_bytes_per_launch"]
            except Exception:
                pass
        if world == 1 and not args.no_cpu_baseline:
            nsamp = min(args.cpu_sample_kf, F)
            sec, cores, kind, build_s, reps = oracle_sample_rate(pack.shard(0, nsamp), X[args.warmup:], min_seconds=8.0)
            cpu_val = (nsamp / F) / sec
            line["cpu_baseline"] = {
                "value": cpu_val, "unit": UNIT, "cores": cores, "kind": "reference" if kind == "ref" else "port",
                "sample": f"{nsamp} of {F} keyframes x {reps} candidates (same step: BAError + association + linearisation), "
                          f"OpenMP over keyframes, extrapolated linearly; KNN = {'reference nanoflann v1.5.0' if kind == 'ref' else 'nanoflann port'}; "
                          f"KD-tree build {build_s:.2f}s excluded",
            }
            # SURVEY §8d asks for both CPU modes of the BAError cost alone: M1 = OpenMP over keyframes, one candidate
            # at a time (iba_func.cpp:203,463); M2 = candidates in parallel, each serial over keyframes (NOMAD threads)
            try:
                from oracle import oracle as O
                n2 = min(96, F)
                o2 = O.Oracle(pack.shard(0, n2), kind="ref" if O.have_ref() else "port", nthreads=cores)
                Xc = X[args.warmup: args.warmup + cores] if len(X) - args.warmup >= cores else np.resize(X[args.warmup:], (cores, 7))
                t0 = time.perf_counter(); o2.ba_error_sums(Xc, mode=1, strict=True, nthreads=cores); t_m1 = time.perf_counter() - t0
                t0 = time.perf_counter(); o2.ba_error_sums(Xc, mode=2, strict=True, nthreads=cores); t_m2 = time.perf_counter() - t0
                line["cpu_baseline"]["bae_only"] = {
                    "m1_keyframe_parallel_evals_per_s": len(Xc) * (n2 / F) / t_m1, "m2_candidate_parallel_evals_per_s": len(Xc) * (n2 / F) / t_m2,
                    "sample": f"BAError only, {len(Xc)} candidates x {n2} of {F} keyframes, {cores} threads, extrapolated linearly"}
            except Exception as e:  # the headline baseline above stands on its own
                line["cpu_baseline"]["bae_only"] = {"error": str(e)[:200]}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
