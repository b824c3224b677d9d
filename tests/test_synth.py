"""Synthetic KeyFramePack generator: determinism, shape, shard consistency."""
import numpy as np


def test_deterministic_and_kitti_shaped(synth):
    a, xa, _ = synth.generate(n_kf=3, seed=7)
    b, xb, _ = synth.generate(n_kf=3, seed=7)
    for name in ("scan_xyz", "kp_xy", "kp_mappoint", "covis_uv", "Tcw", "he_Tl", "covis_relpose"):
        assert np.array_equal(getattr(a, name), getattr(b, name), equal_nan=True), name
    assert np.array_equal(xa, xb)
    n = np.diff(a.scan_offset)
    assert (n > 100_000).all() and (n <= 120_000).all()  # 64 x 1875 rays minus no-returns: ragged
    assert a.n_keypoints == 3 * 2000
    assert a.intrinsics[0].tolist() == [np.float32(718.856), np.float32(718.856), np.float32(607.1928), np.float32(185.2157)]
    assert a.image_wh[0].tolist() == [1241, 376]
    assert a.he_valid.tolist() == [1, 1, 0]
    c, _, _ = synth.generate(n_kf=3, seed=8)
    assert not np.array_equal(a.scan_xyz[:1000], c.scan_xyz[:1000])


def test_shard_is_bitwise_slice_of_full_pack(synth):
    full, _, _ = synth.generate(n_kf=5, n_kf_total=5, seed=3)
    part, _, _ = synth.generate(n_kf=2, kf_begin=2, n_kf_total=5, seed=3)
    ref = full.shard(2, 4)
    for name in ("scan_offset", "scan_xyz", "kp_xy", "kp_mappoint", "covis_uv", "Tcw", "he_Tl", "he_Tc", "covis_relpose",
                 "covis_valid", "he_valid"):
        assert np.array_equal(getattr(part, name), getattr(ref, name), equal_nan=True), name
    assert ref.he_valid.tolist() == [1, 1]  # a successor exists globally, even at the shard's end


def test_no_duplicate_points(synth):
    p, _, _ = synth.generate(n_kf=1, seed=11)
    v = np.ascontiguousarray(p.scan_xyz).view([("", np.float32)] * 3)
    assert len(np.unique(v)) == len(v)
