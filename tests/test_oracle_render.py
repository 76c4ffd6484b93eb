"""Oracle self-checks: analytic cases and an independent fp64 numpy restatement written in a
different style (vectorised einsum), as a second opinion on oracle/splat_oracle.c."""
import numpy as np
import pytest

import oracle
from sim_a_splat_b200 import scene as S

C0 = S.C0


def one_gaussian_scene(mean, scale, opacity, rgb, quat=(1, 0, 0, 0), sh_degree=0):
    K = (sh_degree + 1) ** 2
    sh = np.zeros((1, K, 3), dtype=np.float32)
    sh[0, 0] = (np.asarray(rgb, dtype=np.float64) - 0.5) / C0
    return dict(means=np.array([mean], dtype=np.float32), quats=np.array([quat], dtype=np.float32),
                scales=np.array([[scale] * 3 if np.isscalar(scale) else scale], dtype=np.float32),
                opacities=np.array([opacity], dtype=np.float32), sh=sh, link_id=np.array([-1], dtype=np.int32), L=0)


def cat_scenes(*scenes):
    out = {k: np.concatenate([s[k] for s in scenes]) for k in ("means", "quats", "scales", "opacities", "sh", "link_id")}
    out["L"] = max(s["L"] for s in scenes)
    return out


def test_single_isotropic_gaussian_on_axis():
    W, H, f, z, s, o = 64, 48, 50.0, 2.0, 0.1, 0.8
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1.0]])
    sc = one_gaussian_scene((0, 0, z), s, o, (0.9, 0.4, 0.2))
    fr = oracle.render(sc, np.eye(4)[None], K[None], W, H, bg=(0.1, 0.2, 0.3))
    var = (f * s / z) ** 2 + 0.3
    assert fr.radii[0, 0].tolist() == [int(np.ceil(3.33 * np.sqrt(var)))] * 2
    np.testing.assert_allclose(fr.means2d[0, 0], [W / 2, H / 2], atol=1e-5)
    np.testing.assert_allclose(fr.conics[0, 0], [1 / var, 0, 1 / var], atol=1e-6)
    np.testing.assert_allclose(fr.colors[0, 0], [0.9, 0.4, 0.2], atol=1e-6)
    yy, xx = np.mgrid[0:H, 0:W]
    d2 = (xx + 0.5 - W / 2) ** 2 + (yy + 0.5 - H / 2) ** 2
    al = np.minimum(0.999, o * np.exp(-0.5 * d2 / var))
    al = np.where(al < 1 / 255, 0.0, al)
    np.testing.assert_allclose(fr.alpha[0], al, atol=2e-6)
    want = al[..., None] * np.array([0.9, 0.4, 0.2]) + (1 - al[..., None]) * np.array([0.1, 0.2, 0.3])
    np.testing.assert_allclose(fr.rgb[0], want, atol=3e-6)
    np.testing.assert_allclose(fr.depth[0][al > 0], z, rtol=2e-5)  # 1-T cancels for faint pixels
    assert np.all(fr.depth[0][al == 0] == 0.0)
    assert np.array_equal(fr.n_contrib[0] > 0, al > 0)


def test_two_overlapping_gaussians_front_to_back():
    W, H, f = 32, 32, 40.0
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1.0]])
    near = one_gaussian_scene((0, 0, 1.0), 0.05, 0.6, (1.0, 0.0, 0.0))
    far = one_gaussian_scene((0, 0, 3.0), 0.3, 0.9, (0.0, 0.0, 1.0))
    for sc, order in ((cat_scenes(near, far), [0, 1]), (cat_scenes(far, near), [1, 0])):
        fr = oracle.render(sc, np.eye(4)[None], K[None], W, H)
        cy, cx = H // 2, W // 2
        lo, hi = fr.isect_offsets[0, cy // 16, cx // 16], fr.M
        tile_ids = fr.flatten_ids[lo:hi] if (cy // 16, cx // 16) == (1, 1) else None
        assert tile_ids is not None and tile_ids[:2].tolist() == order  # nearer Gaussian first
        v1, v2 = (f * 0.05 / 1.0) ** 2 + 0.3, (f * 0.3 / 3.0) ** 2 + 0.3
        d2 = 0.5
        a1, a2 = 0.6 * np.exp(-0.5 * d2 / v1), 0.9 * np.exp(-0.5 * d2 / v2)
        want = np.array([a1, 0.0, a2 * (1 - a1)])
        np.testing.assert_allclose(fr.rgb[0, cy, cx], want, atol=3e-6)
        np.testing.assert_allclose(fr.alpha[0, cy, cx], 1 - (1 - a1) * (1 - a2), atol=3e-6)
        ed = (a1 * 1.0 + a2 * (1 - a1) * 3.0) / (1 - (1 - a1) * (1 - a2))
        np.testing.assert_allclose(fr.depth[0, cy, cx], ed, rtol=3e-6)
        assert fr.n_contrib[0, cy, cx] == 2


def test_gaussian_straddling_a_tile_corner():
    W, H, f = 64, 64, 64.0
    K = np.array([[f, 0, 32.0], [0, f, 32.0], [0, 0, 1.0]])
    sc = one_gaussian_scene((0, 0, 1.0), 0.02, 0.9, (0.5, 0.5, 0.5))  # centre (32,32) = corner of 4 tiles
    fr = oracle.render(sc, np.eye(4)[None], K[None], W, H)
    r = int(fr.radii[0, 0, 0])
    assert 1 <= r < 16
    assert fr.tile_rect[0, 0].tolist() == [1, 1, 3, 3] and fr.tiles_per_gauss[0, 0] == 4 and fr.M == 4
    tiles = (fr.isect_ids >> 32) & ((1 << fr.tile_bits) - 1)
    assert tiles.tolist() == [5, 6, 9, 10]  # y*tile_w+x for (1,1),(2,1),(1,2),(2,2), sorted
    assert fr.isect_offsets[0].ravel().tolist() == [0] * 6 + [1, 2, 2, 2, 3] + [4] * 5


def test_empty_view_and_ragged_image_size():
    sc = S.make_synthetic_scene(500, 1, seed=5).as_oracle_dict()
    K = S.fov_intrinsics(50, 37)
    away = S.lookat_viewmat((5.0, 0, 0), (10.0, 0, 0))  # everything is behind the camera
    assert np.all(np.isfinite(away))
    fr = oracle.render(sc, away[None], K[None], 50, 37, bg=(0.2, 0.3, 0.4))
    assert fr.M == 0 and np.all(fr.radii == 0) and np.all(fr.alpha == 0) and np.all(fr.depth == 0)
    np.testing.assert_array_equal(fr.rgb, np.broadcast_to(np.float32([0.2, 0.3, 0.4]), fr.rgb.shape))
    assert (fr.tile_w, fr.tile_h) == (4, 3)
    fr = oracle.render(sc, S.lookat_viewmat((2.0, 1.5, 1.0), (0, 0, 0))[None], K[None], 50, 37)
    assert fr.M > 0 and fr.rgb.shape == (1, 37, 50, 3) and fr.alpha.max() > 0.5
    assert fr.tile_rect[..., 2].max() <= 4 and fr.tile_rect[..., 3].max() <= 3


def test_sorted_keys_offsets_and_tie_break():
    sc = S.make_synthetic_scene(4000, 2, seed=9, link_counts=(300, 200))
    d = sc.as_oracle_dict()
    # exact duplicates -> identical depth keys; stable sort must keep ascending flattened index
    for k in ("means", "quats", "scales", "opacities", "sh", "link_id"):
        d[k] = np.concatenate([d[k], d[k][:500]])
    W, H = 96, 80
    vms = S.ring_cameras(3, (0, 0, 0), seed=2)
    poses = S.random_link_poses(2, 2, seed=77)
    fr = oracle.render(d, vms, S.fov_intrinsics(W, H)[None].repeat(3, 0), W, H, poses, [0, 1, 1])
    assert np.all(np.diff(fr.isect_ids) >= 0)
    same = np.diff(fr.isect_ids) == 0
    assert same.sum() > 50  # the duplicates do tie
    assert np.all(np.diff(fr.flatten_ids.astype(np.int64))[same] > 0)
    assert fr.M == fr.tiles_per_gauss.sum()
    n_tiles = fr.tile_w * fr.tile_h
    key_hi = fr.isect_ids >> 32
    want = np.searchsorted(key_hi, (np.arange(3)[:, None] << fr.tile_bits | np.arange(n_tiles)[None]).ravel(), side="left")
    np.testing.assert_array_equal(fr.isect_offsets.ravel(), want)
    assert fr.tile_bits == int(n_tiles).bit_length() and fr.image_bits == 2
    bits = fr.depths.view(np.uint32).ravel()[fr.flatten_ids].astype(np.int64)
    np.testing.assert_array_equal(fr.isect_ids & 0xFFFFFFFF, bits)
    np.testing.assert_array_equal(fr.isect_ids >> (32 + fr.tile_bits), fr.flatten_ids // d["means"].shape[0])


def test_identity_link_pose_equals_static_render():
    sc = S.make_synthetic_scene(3000, 3, seed=21, link_counts=(200, 150, 90))
    d = sc.as_oracle_dict()
    static = dict(d, link_id=-np.ones(sc.N, dtype=np.int32), L=0)
    ident = np.zeros((1, 3, 12), dtype=np.float32)
    ident[..., 0] = ident[..., 4] = ident[..., 8] = 1.0
    W, H = 80, 64
    vm, K = S.ring_cameras(1, (0, 0, 0), seed=4), S.fov_intrinsics(W, H)[None]
    a = oracle.render(d, vm, K, W, H, ident, [0])
    b = oracle.render(static, vm, K, W, H)
    for k in ("radii", "means2d", "depths", "conics", "colors", "isect_ids", "flatten_ids", "rgb", "depth", "alpha", "n_contrib"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    moved = oracle.render(d, vm, K, W, H, S.random_link_poses(1, 3, 5), [0])
    assert np.abs(moved.rgb - a.rgb).max() > 1e-3


# --- independent fp64 restatement (vectorised; deliberately not sharing code or op order) ---------

def _sh_basis64(deg, d):
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    b = [np.full_like(x, 0.2820947917738781)]
    if deg >= 1:
        b += [-0.48860251190292 * y, 0.48860251190292 * z, -0.48860251190292 * x]
    if deg >= 2:
        b += [1.092548430592079 * x * y, -1.092548430592079 * y * z, 0.9461746957575601 * z * z - 0.3153915652525201,
              -1.092548430592079 * x * z, 0.5462742152960395 * (x * x - y * y)]
    if deg >= 3:
        b += [-0.5900435899266435 * y * (3 * x * x - y * y), 2.890611442640554 * x * y * z,
              -0.4570457994644658 * y * (5 * z * z - 1), 0.3731763325901154 * z * (5 * z * z - 3),
              -0.4570457994644658 * x * (5 * z * z - 1), 1.445305721320277 * z * (x * x - y * y),
              -0.5900435899266435 * x * (x * x - 3 * y * y)]
    return np.stack(b, axis=1)


def render64(d, vm, K, W, H, poses=None, bg=(0, 0, 0)):
    mu, q, s = (np.asarray(d[k], dtype=np.float64) for k in ("means", "quats", "scales"))
    op, sh, link = np.asarray(d["opacities"], np.float64), np.asarray(d["sh"], np.float64), d["link_id"]
    N = mu.shape[0]
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z),
                  2 * (y * z - w * x), 2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], axis=1).reshape(N, 3, 3)
    rot = np.tile(np.eye(3), (N, 1, 1))
    pos = np.zeros((N, 3))
    if poses is not None:
        m = link >= 0
        P = np.asarray(poses, np.float64)[0][link[m]]
        rot[m], pos[m] = P[:, :9].reshape(-1, 3, 3), P[:, 9:]
    mu = np.einsum("nij,nj->ni", rot, mu) + pos
    R = rot @ R
    Sig = np.einsum("nij,nj,nkj->nik", R, s * s, R)
    Rc, tc = vm[:3, :3], vm[:3, 3]
    mc = mu @ Rc.T + tc
    Cc = Rc @ Sig @ Rc.T
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    zc = mc[:, 2]
    ok = (zc >= 0.01) & (zc <= 1e10)
    zs = np.where(ok, zc, 1.0)
    tfx, tfy = 0.5 * W / fx, 0.5 * H / fy
    tx = zs * np.clip(mc[:, 0] / zs, -(cx / fx + 0.3 * tfx), (W - cx) / fx + 0.3 * tfx)
    ty = zs * np.clip(mc[:, 1] / zs, -(cy / fy + 0.3 * tfy), (H - cy) / fy + 0.3 * tfy)
    J = np.zeros((N, 2, 3))
    J[:, 0, 0], J[:, 0, 2], J[:, 1, 1], J[:, 1, 2] = fx / zs, -fx * tx / zs**2, fy / zs, -fy * ty / zs**2
    c2 = J @ Cc @ J.transpose(0, 2, 1)
    c2[:, 0, 0] += 0.3
    c2[:, 1, 1] += 0.3
    det = c2[:, 0, 0] * c2[:, 1, 1] - c2[:, 0, 1] ** 2
    m2 = np.stack([fx * mc[:, 0] / zs + cx, fy * mc[:, 1] / zs + cy], axis=1)
    rad = np.ceil(3.33 * np.sqrt(np.stack([c2[:, 0, 0], c2[:, 1, 1]], axis=1)))
    ok &= det > 0
    ok &= (m2[:, 0] + rad[:, 0] > 0) & (m2[:, 0] - rad[:, 0] < W) & (m2[:, 1] + rad[:, 1] > 0) & (m2[:, 1] - rad[:, 1] < H)
    conic = np.stack([c2[:, 1, 1] / det, -c2[:, 0, 1] / det, c2[:, 0, 0] / det], axis=1)
    campos = -Rc.T @ tc
    dirs = np.einsum("nji,nj->ni", rot, mu - campos)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    deg = int(round(np.sqrt(sh.shape[1]))) - 1
    col = np.maximum(np.einsum("nk,nkc->nc", _sh_basis64(deg, dirs), sh) + 0.5, 0.0)
    order = np.argsort(np.where(ok, zc, np.inf), kind="stable")
    order = order[: int(ok.sum())]
    yy, xx = np.mgrid[0:H, 0:W]
    px, py = xx + 0.5, yy + 0.5
    T = np.ones((H, W))
    acc = np.zeros((H, W, 3))
    accd = np.zeros((H, W))
    live = np.ones((H, W), dtype=bool)
    for g in order:
        dx, dy = m2[g, 0] - px, m2[g, 1] - py
        inbox = (np.abs(dx) <= rad[g, 0] + 16) & (np.abs(dy) <= rad[g, 1] + 16)  # same-tile reach; alpha cut decides anyway
        sig = 0.5 * (conic[g, 0] * dx * dx + conic[g, 2] * dy * dy) + conic[g, 1] * dx * dy
        al = np.minimum(0.999, op[g] * np.exp(-sig))
        use = live & inbox & (sig >= 0) & (al >= 1 / 255)
        nT = T * (1 - al)
        stop = use & (nT <= 1e-4)
        live &= ~stop
        use &= ~stop
        vis = np.where(use, al * T, 0.0)
        acc += vis[..., None] * col[g]
        accd += vis * zc[g]
        T = np.where(use, nT, T)
    A = 1 - T
    return dict(ok=ok, means2d=m2, radii=rad, conics=conic, colors=col, depths=zc, rgb=acc + T[..., None] * np.asarray(bg),
                alpha=A, depth=accd / np.maximum(A, 1e-10))


@pytest.mark.parametrize("deg,seed", [(0, 1), (1, 2), (2, 3), (3, 4)])
def test_oracle_against_independent_fp64(deg, seed):
    sc = S.make_synthetic_scene(1500, deg, seed=seed, link_counts=(120, 80))
    d = sc.as_oracle_dict()
    d["scales"] = d["scales"] * 3.0  # bigger splats: more overlap per pixel
    W, H = 72, 56
    vm = S.ring_cameras(1, (0, 0, 0), seed=seed)[0]
    K = S.fov_intrinsics(W, H)
    poses = S.random_link_poses(1, 2, seed=seed + 10)
    fr = oracle.render(d, vm[None], K[None], W, H, poses, [0], bg=(0.3, 0.2, 0.1))
    ref = render64(d, vm.astype(np.float32).astype(np.float64), K.astype(np.float32).astype(np.float64), W, H, poses, bg=(0.3, 0.2, 0.1))
    vis = fr.radii[0, :, 0] > 0
    assert np.mean(vis != ref["ok"]) < 0.002
    both = vis & ref["ok"]
    np.testing.assert_allclose(fr.means2d[0][both], ref["means2d"][both], rtol=2e-5, atol=2e-3)
    np.testing.assert_allclose(fr.conics[0][both], ref["conics"][both], rtol=2e-3, atol=1e-5)
    np.testing.assert_allclose(fr.colors[0][both], ref["colors"][both], atol=2e-5)
    np.testing.assert_allclose(fr.depths[0][both], ref["depths"][both], rtol=1e-5)
    assert np.mean(np.any(fr.radii[0][both] != ref["radii"][both], axis=1)) < 0.005
    # images: fp32 vs fp64 decisions differ on a handful of borderline alphas only
    assert np.abs(fr.rgb[0] - ref["rgb"]).max() < 5e-3 and np.abs(fr.rgb[0] - ref["rgb"]).mean() < 2e-5
    assert np.abs(fr.alpha[0] - ref["alpha"]).max() < 5e-3
    m = ref["alpha"] > 0.05
    assert np.abs(fr.depth[0][m] - ref["depth"][m]).max() < 5e-3


def test_radius_rule_cuda_1_5_on_the_oracle():
    """The optional gsplat-1.5-CUDA radius rule (opacity-aware extend, culls opacity < 1/255): smaller radii, never
    larger; the same images as the pinned default (a pixel outside either box has alpha < 1/255)."""
    from sim_a_splat_b200 import scene as S

    sc = S.make_synthetic_scene(4000, 2, seed=3, link_counts=(200,))
    sc.opacities[::7] = np.float32(0.003)          # below 1/255: culled by the rule
    sc.opacities[1::7] = np.float32(0.02)
    cams = S.ring_cameras(2, (0, 0, 0), seed=5)
    K = np.repeat(S.fov_intrinsics(96, 80)[None], 2, 0)
    poses = S.random_link_poses(1, 1, seed=2)
    a = oracle.render(sc.as_oracle_dict(), cams, K, 96, 80, poses, [0, 0])
    b = oracle.render(sc.as_oracle_dict(), cams, K, 96, 80, poses, [0, 0], radius_rule="cuda_1_5")
    assert np.all(b["radii"] <= a["radii"]) and b["M"] < a["M"]
    assert np.all(b["radii"][:, ::7] == 0)
    for k in ("rgb", "depth", "alpha", "n_contrib"):
        np.testing.assert_array_equal(a[k], b[k])
    lib = oracle.oracle._lib()
    assert lib.orc_extend_export(0, 0.5) == np.float32(3.33) and lib.orc_extend_export(1, 0.003) == 0.0
    assert abs(lib.orc_extend_export(1, 0.02) - np.sqrt(2 * np.log(0.02 * 255))) < 1e-6
    assert abs(lib.orc_extend_export(1, 0.99) - np.sqrt(2 * np.log(0.99 * 255))) < 1e-6   # 3.326: even opaque Gaussians shrink a little
    assert lib.orc_extend_export(1, 1.5) == np.float32(3.33)                              # the cap
