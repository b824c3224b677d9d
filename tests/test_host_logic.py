"""Host-side logic that needs no GPU: keyframe sharding, additivity of the partial sums, and the
world_size-2 all-reduce (gloo) that the multi-GPU path uses."""
import importlib
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from conftest import PKG, ROOT


def test_shard_bounds_partition():
    par = importlib.import_module(PKG + ".parallel")
    for n in (1, 7, 50, 1500):
        for w in (1, 2, 3, 8):
            b = [par.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1


def test_partial_sums_are_additive_over_keyframe_shards(oracle_mod, small_pack, small_candidates):
    """Sharding by keyframe changes no decision (all gates are per keyframe)."""
    par = importlib.import_module(PKG + ".parallel")
    pack, _ = small_pack
    full, _, _ = oracle_mod.Oracle(pack).ba_error_sums(small_candidates, mode=1)
    acc = np.zeros_like(full)
    for r in range(3):
        b, e = par.shard_bounds(pack.n_kf, 3, r)
        s, _, _ = oracle_mod.Oracle(pack.shard(b, e)).ba_error_sums(small_candidates, mode=1)
        acc += s
    assert np.array_equal(acc[:, 3:], full[:, 3:])                 # every counter exactly
    assert np.allclose(acc[:, :3], full[:, :3], rtol=1e-13, atol=0)  # fp64 sums re-associated


_WORKER = textwrap.dedent("""
    import importlib, os, sys
    import numpy as np
    sys.path.insert(0, {root!r})
    import torch.distributed as dist
    PKG = {pkg!r}
    par = importlib.import_module(PKG + ".parallel")
    synth = importlib.import_module(PKG + ".synth")
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    rank, world = dist.get_rank(), dist.get_world_size()
    F = 4
    b, e = par.shard_bounds(F, world, rank)
    # each rank generates ITS keyframes only (bitwise the slice of the full pack)
    pack, x_gt, _ = synth.generate(n_kf=e - b, kf_begin=b, n_kf_total=F, beams=16, az_steps=600, n_kp=300, seed=5)
    X = synth.candidates(x_gt, 3, 0.3)
    part, _, _ = O.Oracle(pack).ba_error_sums(X, mode=0)
    tot = par.SumAllReduce()(part)
    if rank == 0:
        full_pack, _, _ = synth.generate(n_kf=F, seed=5, beams=16, az_steps=600, n_kp=300)
        full, _, _ = O.Oracle(full_pack).ba_error_sums(X, mode=0)
        assert np.array_equal(tot[:, 3:], full[:, 3:]), (tot, full)
        assert np.allclose(tot[:, :3], full[:, :3], rtol=1e-13, atol=0)
        print("OK")
    dist.destroy_process_group()
""")


def test_world_size_2_allreduce_over_gloo(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, pkg=PKG, port=port))
    env = dict(os.environ, OMP_NUM_THREADS="2")
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
             for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK" in outs[0][0]


def test_pack_npz_roundtrip(pkg, small_pack):
    pack, _ = small_pack
    sub = pack.shard(1, 3)
    back = pkg.KeyFramePack.from_npz_dict(sub.to_npz_dict())
    assert back.n_kf == 2 and np.array_equal(back.scan_xyz, sub.scan_xyz) and np.array_equal(back.he_Tl, sub.he_Tl)
    c = sub.as_c()
    assert c.n_kf == 2 and c.scan_offset[2] == sub.n_points


def test_stand_in_optimisers_on_analytic_problems(pkg):
    """optim.py (the NOMAD / Ceres stand-ins used by the final-extrinsic parity tests) on closed forms."""
    optim = importlib.import_module(PKG + ".optim")
    target = np.array([0.02, -0.03, 0.01, 0.1, -0.2, 0.05, 0.3])

    class Quad:
        def eval_block(self, X):
            X = np.atleast_2d(X)
            # objective + one active constraint x[6] <= 0.25 + two slack ones
            return [[float(((x - target) ** 2).sum()), x[6] - 0.25, -1.0, -1.0] for x in X]

    x, bbo, n_eval, hist = optim.poll_search(Quad(), np.zeros(7), -np.ones(7), np.ones(7), max_bb_eval=1500, init_frame=0.5 * np.ones(7))
    want = target.copy(); want[6] = 0.25
    assert np.abs(x - want).max() < 1e-3 and bbo[1] <= 0 and n_eval <= 1500
    assert all(b[1] <= a[1] or b[2] < a[2] for a, b in zip(hist, hist[1:]))

    class Rosen:
        def build(self, x):
            return None

        def evaluate(self, x):  # residuals r = (10 (x1 - x0^2), 1 - x0) padded to 7 parameters
            r = np.concatenate([[10 * (x[1] - x[0] ** 2), 1 - x[0]], x[2:]])
            J = np.zeros((7, 7)); J[0, 0] = -20 * x[0]; J[0, 1] = 10; J[1, 0] = -1; J[2:, 2:] = np.eye(5)
            return 0.5 * float(r @ r), J.T @ r, J.T @ J

    x, costs = optim.lm_refine(Rosen(), np.array([-1.2, 1.0, 0.3, 0, 0, 0, 0.1]), max_iba_iter=3, max_num_iterations=60)
    assert np.abs(x[:2] - 1).max() < 1e-6 and np.abs(x[2:]).max() < 1e-8 and costs[-1] < 1e-12
