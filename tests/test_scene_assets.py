"""Asset ingestion: the canonical order reproduces the reference's partition
(splat_handler.py:104-143) on the REAL masks shipped with the reference."""
import numpy as np

from sim_a_splat_b200 import scene as S


def test_real_mask_counts(assets):
    m = S.unpack_masks(assets, "robots-scene-v2")
    assert m.shape == (7, 113831)
    assert m.sum(axis=1).tolist() == [534, 467, 1387, 751, 424, 302, 89]
    assert m.any(axis=0).sum() == 3448 and (m.sum(axis=0) > 1).sum() == 499
    d = S.unpack_masks(assets, "divar113vhw")
    assert d.shape == (6, 292247)
    assert d.sum(axis=1).tolist() == [3342, 15563, 10528, 7372, 7365, 6219]
    assert d.any(axis=0).sum() == 44048


def test_canonical_order_duplicates_overlaps(assets):
    sc = S.make_reference_like_scene("robots-scene-v2", sh_degree=1, seed=3, assets=assets)
    m = S.unpack_masks(assets, "robots-scene-v2")
    assert sc.N == 110383 + 3954 and sc.L == 7 and sc.n_background == 110383
    assert np.all(np.diff(sc.orig_index[: sc.n_background]) > 0)
    off = sc.n_background
    for l in range(7):
        idx = np.nonzero(m[l])[0]
        np.testing.assert_array_equal(sc.orig_index[off: off + idx.size], idx)
        assert np.all(sc.link_id[off: off + idx.size] == l)
        off += idx.size
    assert off == sc.N
    # a Gaussian claimed by two masks appears twice with identical parameters
    dup = np.nonzero(m.sum(axis=0) == 2)[0][0]
    where = np.nonzero(sc.orig_index == dup)[0]
    assert where.size == 2 and np.array_equal(sc.means[where[0]], sc.means[where[1]])


def test_tblock_segment_and_determinism(assets):
    a = S.make_reference_like_scene("robots-scene-v2", 3, 101, with_tblock=True, assets=assets)
    b = S.make_reference_like_scene("robots-scene-v2", 3, 101, with_tblock=True, assets=assets)
    assert a.N == 114337 + 4000 and a.L == 8 and np.sum(a.link_id == 7) == 4000
    for k in ("means", "quats", "scales", "opacities", "sh", "link_id"):
        np.testing.assert_array_equal(getattr(a, k), getattr(b, k))
    assert a.sh.shape == (a.N, 16, 3)
    assert 0.1 <= a.opacities.min() and a.opacities.max() <= 0.995
    assert 0.001 - 1e-6 <= a.scales.min() and a.scales.max() <= 0.08 + 1e-6


def test_camera_conventions():
    vm = S.lookat_viewmat((1.0, 2.0, 3.0), (0, 0, 0))
    p = vm @ np.array([0, 0, 0, 1.0])
    assert abs(p[0]) < 1e-12 and abs(p[1]) < 1e-12 and abs(p[2] - np.sqrt(14)) < 1e-12  # target on +z
    np.testing.assert_allclose(vm[:3, :3] @ vm[:3, :3].T, np.eye(3), atol=1e-12)
    # viser-side pose (c2w, OpenCV) and nerfstudio-side pose (c2w, OpenGL) describe the same camera
    c2w = np.linalg.inv(vm)
    from sim_a_splat_b200.transforms import SO3

    np.testing.assert_allclose(S.c2w_opencv_to_viewmat(SO3.from_matrix(c2w[:3, :3]).wxyz, c2w[:3, 3]), vm, atol=1e-12)
    gl = c2w.copy()
    gl[:3, 1:3] *= -1
    np.testing.assert_allclose(S.c2w_opengl_to_viewmat(gl[:3]), vm, atol=1e-12)
