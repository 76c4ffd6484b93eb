"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Bar (BASELINE.json north_star): tile ids, sort order, per-pixel contributor
counts bit-exact; RGB/depth max-abs <= 1e-3 with PSNR > 60 dB."""
import numpy as np
import pytest
import torch

import oracle
from sim_a_splat_b200 import scene as S
from sim_a_splat_b200.renderer import NS_BACKGROUND, PipelinedSplatRenderer, SplatRenderer

from parity_utils import assert_parity, compare, run_both

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_small_scene_all_stages(deg):
    sc = S.make_synthetic_scene(6000, deg, seed=40 + deg, link_counts=(400, 250, 100))
    W, H = 112, 80
    vms = S.ring_cameras(4, (0, 0, 0), seed=deg)
    poses = S.random_link_poses(3, 3, seed=500 + deg)
    gpu, ref = run_both(sc, vms, S.fov_intrinsics(W, H)[None], W, H, poses, [0, 1, 2, 1], bg=(0.1, 0.5, 0.9))
    rep = compare(gpu, ref)
    print(rep)
    assert_parity(rep)
    assert rep["M_ref"] > 1000


def test_views_shared_across_envs_and_mixed_cameras():
    """2 fixed cameras x 5 envs = 2 views / 10 images, plus one env-specific (moving) camera each:
    background work is shared per view, results must not notice."""
    sc = S.make_synthetic_scene(9000, 2, seed=61, link_counts=(700, 300))
    W, H, E = 128, 96, 5
    fixed = S.ring_cameras(2, (0, 0, 0), seed=1)
    moving = S.ring_cameras(E, (0, 0, 0), seed=2)
    vms, env = [], []
    for e in range(E):
        vms += [fixed[0], moving[e], fixed[1]]
        env += [e, e, e]
    gpu, ref = run_both(sc, np.stack(vms), S.fov_intrinsics(W, H)[None], W, H, S.random_link_poses(E, 2, seed=70), env)
    assert gpu["stats"]["n_views"] == 2 + E and gpu["stats"]["n_images"] == 3 * E
    assert gpu["view_of_image"].tolist() == [0, 1, 2, 0, 3, 2, 0, 4, 2, 0, 5, 2, 0, 6, 2]
    rep = compare(gpu, ref)
    print(rep)
    assert_parity(rep)


def test_no_links_static_scene():
    sc = S.make_synthetic_scene(5000, 3, seed=7)
    gpu, ref = run_both(sc, S.ring_cameras(2, (0, 0, 0), seed=3), S.fov_intrinsics(96, 96)[None], 96, 96)
    assert_parity(compare(gpu, ref))


def test_config1_tblock_scene_640x480(assets):
    """BASELINE.json configs[0]: robots-scene-v2 statistics + T-block segment, one 640x480 frame."""
    sc = S.make_reference_like_scene("robots-scene-v2", 3, 101, with_tblock=True, assets=assets)
    cen = sc.means[sc.link_id >= 0].mean(0)
    gpu, ref = run_both(sc, S.ring_cameras(1, cen, 7), S.fov_intrinsics(640, 480)[None], 640, 480,
                        S.random_link_poses(1, sc.L, 1000), [0])
    rep = compare(gpu, ref)
    print(rep)
    assert_parity(rep)


def test_config2_pusht_batched_envs_slice(assets):
    """BASELINE.json configs[1] at 16 of the 256 envs: 96x96, one fixed camera, per-env link poses."""
    sc = S.make_reference_like_scene("robots-scene-v2", 3, 102, with_tblock=True, assets=assets)
    cen = sc.means[sc.link_id >= 0].mean(0)
    E = 16
    vms = np.repeat(S.ring_cameras(1, cen, 8), E, axis=0)
    gpu, ref = run_both(sc, vms, S.fov_intrinsics(96, 96)[None], 96, 96, S.random_link_poses(E, sc.L, 1000), np.arange(E))
    rep = compare(gpu, ref)
    print(rep)
    assert_parity(rep)


def test_config3_xarm_real_masks_720p(assets):
    """BASELINE.json configs[2]: real xarm6-1 masks (L=7), 2 of the 4 cameras at 1280x720."""
    sc = S.make_reference_like_scene("robots-scene-v2", 3, 103, assets=assets)
    cen = sc.means[sc.link_id >= 0].mean(0)
    gpu, ref = run_both(sc, S.ring_cameras(2, cen, 9), S.fov_intrinsics(1280, 720)[None], 1280, 720,
                        S.random_link_poses(1, 7, 2000), [0, 0])
    rep = compare(gpu, ref)
    print(rep)
    assert_parity(rep)


def test_config4_divar_slice(assets):
    """BASELINE.json configs[3] slice: divar113vhw real masks (L=6), 3 envs x 2 cameras at 224x224."""
    sc = S.make_reference_like_scene("divar113vhw", 3, 104, assets=assets)
    cen = sc.means[sc.link_id >= 0].mean(0)
    cams = S.ring_cameras(2, cen, 10)
    vms = np.concatenate([cams] * 3)
    gpu, ref = run_both(sc, vms, S.fov_intrinsics(224, 224)[None], 224, 224, S.random_link_poses(3, 6, 3000), [0, 0, 1, 1, 2, 2])
    rep = compare(gpu, ref)
    print(rep)
    assert_parity(rep)


def test_empty_view_ragged_size_and_single_gaussian():
    sc = S.make_synthetic_scene(500, 1, seed=5)
    K = S.fov_intrinsics(50, 37)[None]
    gpu, ref = run_both(sc, S.lookat_viewmat((5.0, 0, 0), (10.0, 0, 0))[None], K, 50, 37, bg=(0.2, 0.3, 0.4))
    assert ref["M"] == 0
    assert_parity(compare(gpu, ref))
    gpu, ref = run_both(sc, S.lookat_viewmat((2.0, 1.5, 1.0), (0, 0, 0))[None], K, 50, 37)
    assert_parity(compare(gpu, ref))
    one = S.canonicalize(np.float32([[0, 0, 2.0]]), np.float32([[1, 0, 0, 0]]), np.float32([[0.1, 0.1, 0.1]]), np.float32([0.8]),
                         np.float32([[[1.0, -0.3, 0.2]]]), None)
    Kc = np.float32([[50, 0, 32], [0, 50, 24], [0, 0, 1]])[None]
    gpu, ref = run_both(one, np.eye(4)[None], Kc, 64, 48, bg=(0.1, 0.2, 0.3))
    assert_parity(compare(gpu, ref))


def _cluster_scene(sizes, seed, spread=0.004, opacity=0.02):
    """Dense clusters of tiny, faint Gaussians, one per image tile -> long per-tile lists whose
    compositing does not terminate early (exercises every bucket-sort size class)."""
    rng = np.random.default_rng(seed)
    means, n = [], 0
    for i, sz in enumerate(sizes):
        c = np.array([-0.6 + 0.4 * i, 0.0, 0.0])
        means.append(c + rng.normal(0, spread, size=(sz, 3)))
        n += sz
    means = np.concatenate(means).astype(np.float32)
    quats = rng.normal(size=(n, 4)).astype(np.float32)
    scales = np.full((n, 3), 0.002, dtype=np.float32)
    op = np.full(n, opacity, dtype=np.float32)
    sh = (rng.uniform(0.1, 0.9, size=(n, 1, 3)).astype(np.float32) - 0.5) / S.C0
    # exact duplicates of the first 300 -> depth ties inside a big bucket
    dup = slice(0, 300)
    return S.canonicalize(np.concatenate([means, means[dup]]), np.concatenate([quats, quats[dup]]),
                          np.concatenate([scales, scales[dup]]), np.concatenate([op, op[dup]]),
                          np.concatenate([sh, sh[dup]]), None)


def test_bucket_sort_size_classes_and_depth_ties():
    sc = _cluster_scene((3000, 7000, 10000, 16000), seed=11)
    W, H = 256, 64
    vm = S.lookat_viewmat((0.0, -2.5, 0.0), (0.0, 0.0, 0.0))
    gpu, ref = run_both(sc, vm[None], S.fov_intrinsics(W, H, 40.0)[None], W, H)
    counts = np.diff(np.append(ref["isect_offsets"].ravel(), ref["M"]))
    assert counts.max() > 12288 and np.any((counts > 6144) & (counts <= 12288)) and np.any((counts > 2048) & (counts <= 6144))
    assert np.sum(np.diff(ref["isect_ids"]) == 0) >= 100  # ties present
    rep = compare(gpu, ref)
    print(rep, counts.max())
    assert_parity(rep)
    assert gpu["stats"]["max_tile_list"] == counts.max()


def test_giant_buckets_chunk_sort_and_merge_passes():
    """Buckets far beyond shared memory (round 2: chunks sorted in shared memory by all SMs + multi-CTA merge-path passes,
    csrc/sas_sort.cu big_*): tile lists of 85 000 and 205 300 entries = 7 and 17 chunks (3 and 5 merge passes, odd run
    counts, both ping-pong parities), half of the entries in equal-depth runs that straddle chunks -- the merge orders
    ties by Gaussian id."""
    sizes = (30000, 55000, 150000)
    sc = _cluster_scene(sizes, seed=12, spread=0.003, opacity=0.004)
    # more ties, far apart in emission order: the last 2000 Gaussians are exact copies of 2000 of the big cluster
    n0 = sizes[0] + sizes[1]
    for arr in (sc.means, sc.quats, sc.scales, sc.opacities, sc.sh):
        arr[-2000:] = arr[n0 + 1000:n0 + 3000]
    W, H = 192, 48
    vm = S.lookat_viewmat((0.0, -2.5, 0.0), (-0.2, 0.0, 0.0))
    K = S.fov_intrinsics(W, H, 40.0)
    r = SplatRenderer(sc, max_images=1, max_width=W, max_height=H, max_isects=2_000_000)
    # the path's launches are only enqueued once the renderer has SEEN such lists (a host-side hint from the counters):
    # the first checked render goes through the one-CTA kernel and sets it ...
    r.set_cameras(vm[None], K, W, H)
    first = {k: v.clone() for k, v in r.render(("rgb", "n_contrib")).items()}
    n0 = r.last_launch_count()
    # ... the second one (what run_both compares with the oracle) takes the chunk sort + merge passes
    gpu, ref = run_both(sc, vm[None], K[None], W, H, renderer=r)
    assert r.last_launch_count() == n0 + 8           # plan + chunk sort + 6 merge passes
    counts = np.diff(np.append(ref["isect_offsets"].ravel(), ref["M"]))
    assert counts.max() > 12 * 12288 and np.sum(counts > 4 * 12288) >= 2
    assert np.sum(np.diff(ref["isect_ids"]) == 0) >= 2000
    rep = compare(gpu, ref)
    print(rep, sorted(counts)[-4:])
    assert_parity(rep)
    assert gpu["stats"]["max_tile_list"] == counts.max()
    np.testing.assert_array_equal(first["n_contrib"].cpu().numpy(), gpu["n_contrib"])   # both sort paths agree
    np.testing.assert_array_equal(first["rgb"].cpu().numpy(), gpu["rgb"])
    r.close()


def test_rendering_is_deterministic_and_host_path_matches():
    sc = S.make_synthetic_scene(20000, 2, seed=13, link_counts=(900, 500))
    W, H = 160, 128
    vms = S.ring_cameras(3, (0, 0, 0), seed=5)
    r = SplatRenderer(sc, max_images=3, max_envs=2, max_width=W, max_height=H)
    r.set_link_poses(S.random_link_poses(2, 2, seed=9))
    r.set_cameras(vms, S.fov_intrinsics(W, H), W, H, [0, 1, 0])
    names = ("rgb", "depth", "alpha", "rgb_u8", "n_contrib")
    a = {k: v.clone() for k, v in r.render(names, background=(0.3, 0.3, 0.3)).items()}
    ids_a = r.debug_intermediates()["flatten_ids"]
    b = r.render(names, background=(0.3, 0.3, 0.3))
    for k in names:
        assert torch.equal(a[k], b[k]), k
    np.testing.assert_array_equal(ids_a, r.debug_intermediates()["flatten_ids"])
    h = r.render(names, background=(0.3, 0.3, 0.3), to_host=True)
    for k in names:
        assert not h[k].is_cuda and torch.equal(a[k].cpu(), h[k]), k
    # device-resident link poses (the on-device pose source of SURVEY 8f-1) give the same frame
    r.set_link_poses(torch.from_numpy(S.random_link_poses(2, 2, seed=9)).cuda())
    c = r.render(names, background=(0.3, 0.3, 0.3))
    for k in names:
        assert torch.equal(a[k], c[k]), k
    # linearity in the background: rgb(bg) - rgb(0) = T * bg
    z = r.render(("rgb", "alpha"))
    want = np.broadcast_to(((1 - z["alpha"])[..., None] * 0.3).cpu().numpy(), tuple(a["rgb"].shape))
    np.testing.assert_allclose((a["rgb"] - z["rgb"]).cpu().numpy(), want, atol=2e-7)
    r.close()


def test_cuda_graph_replay_equals_direct_launches():
    """sas_render replays a CUDA graph of the step on real streams; per-stage timing (events between the stages)
    and the legacy default stream take the direct-launch path.  All must agree bit for bit, also after the
    inputs, the outputs or the image count change (graph invalidation)."""
    sc = S.make_synthetic_scene(15000, 2, seed=29, link_counts=(800, 400))
    W, H = 144, 112
    vms = S.ring_cameras(4, (0, 0, 0), seed=8)
    r = SplatRenderer(sc, max_images=4, max_envs=3, max_width=W, max_height=H)
    names = ("rgb", "depth", "alpha", "n_contrib")
    r.set_cameras(vms, S.fov_intrinsics(W, H), W, H, [0, 1, 2, 1])
    poses_a, poses_b = S.random_link_poses(3, 2, seed=1), S.random_link_poses(3, 2, seed=2)
    got = {}
    for tag, poses in (("a", poses_a), ("b", poses_b), ("a2", poses_a)):
        r.set_link_poses(poses)
        got[tag] = {k: v.clone() for k, v in r.render(names).items()}      # own stream -> graph (captured once, replayed)
    for k in names:
        assert torch.equal(got["a"][k], got["a2"][k]), k
        assert not torch.equal(got["a"]["rgb"], got["b"]["rgb"])
    r.set_stage_timing(True)                                                 # direct launches
    r.set_link_poses(poses_b)
    direct = r.render(names)
    assert set(r.stage_ms()) == set(r.STAGES)
    r.set_stage_timing(False)
    for k in names:
        assert torch.equal(direct[k], got["b"][k]), k
    s2 = torch.cuda.Stream()
    with torch.cuda.stream(s2):                                              # caller's own stream -> graph on that stream
        r.set_link_poses(torch.from_numpy(poses_b).cuda())
        on_s2 = r.render(names)
    s2.synchronize()
    for k in names:
        assert torch.equal(on_s2[k], got["b"][k]), k
    r.set_cameras(vms[:2], S.fov_intrinsics(W, H), W, H, [0, 1])             # fewer images: the graph is rebuilt
    r.set_link_poses(poses_a)
    two = r.render(names)
    for k in names:
        assert torch.equal(two[k], got["a"][k][:2]), k
    r.close()


@pytest.mark.gpu
def test_async_host_outputs_pipeline_matches_blocking():
    """render_async / wait (sas_render mem=HOST_ASYNC + sas_wait_outputs): two renders in flight, alternating staging
    sets and two cached graphs.  Every step's host observations must equal the blocking path's, bit for bit, also when
    the wait comes late (after the next render has been issued)."""
    sc = S.make_synthetic_scene(15000, 2, seed=31, link_counts=(800, 400))
    W, H = 112, 96
    E = 6
    vms = S.ring_cameras(1, (0, 0, 0), seed=4)
    r = SplatRenderer(sc, max_images=E, max_envs=E, max_width=W, max_height=H)
    r.set_cameras(np.repeat(vms, E, axis=0), S.fov_intrinsics(W, H), W, H, np.arange(E))
    names = ("rgb_u8", "depth", "n_contrib")
    poses = [S.random_link_poses(E, 2, seed=40 + k) for k in range(7)]
    want = []
    for p in poses:
        r.set_link_poses(p)
        want.append({k: v.clone() for k, v in r.render(names, to_host=True).items()})
    assert not torch.equal(want[0]["rgb_u8"], want[1]["rgb_u8"])
    with pytest.raises(RuntimeError):
        r.wait()                                       # nothing issued yet
    prev = None
    for i, p in enumerate(poses):
        r.set_link_poses(torch.from_numpy(p).pin_memory().numpy() if i % 2 else p)
        tk = r.render_async(names)
        if prev is not None:
            got = r.wait(prev)                         # step i-1 while step i is in flight
            for k in names:
                assert not got[k].is_cuda and torch.equal(got[k], want[i - 1][k]), (i - 1, k)
        prev = tk
    got = r.wait()
    for k in names:
        assert torch.equal(got[k], want[-1][k]), k
    with pytest.raises(RuntimeError):
        r.wait(prev - 2)                               # buffers already reused
    # the blocking path still works afterwards and agrees
    r.set_link_poses(poses[2])
    again = r.render(names, to_host=True)
    for k in names:
        assert torch.equal(again[k], want[2][k]), k
    r.close()


def test_pipelined_renderer_two_steps_in_flight_matches_blocking():
    """PipelinedSplatRenderer: two handles on two streams used alternately, steps overlap on the device.  Every step's
    observations equal the single-handle blocking render of the same poses, bit for bit."""
    sc = S.make_synthetic_scene(15000, 2, seed=33, link_counts=(800, 400))
    W, H, E = 112, 96, 5
    vms = np.repeat(S.ring_cameras(1, (0, 0, 0), seed=4), E, axis=0)
    names = ("rgb_u8", "depth", "n_contrib")
    poses = [S.random_link_poses(E, 2, seed=60 + k) for k in range(9)]
    r = SplatRenderer(sc, max_images=E, max_envs=E, max_width=W, max_height=H)
    r.set_cameras(vms, S.fov_intrinsics(W, H), W, H, np.arange(E))
    want = []
    for p in poses:
        r.set_link_poses(p)
        want.append({k: v.clone() for k, v in r.render(names, to_host=True).items()})
    r.close()
    for depth in (1, 2, 3):
        pr = PipelinedSplatRenderer(sc, depth=depth, max_images=E, max_envs=E, max_width=W, max_height=H)
        pr.set_cameras(vms, S.fov_intrinsics(W, H), W, H, np.arange(E))
        pending = []
        for i, p in enumerate(poses):
            pr.set_link_poses(torch.from_numpy(p).cuda() if i % 3 == 0 else p)
            pending.append((i, pr.render_async(names)))
            if len(pending) > depth:
                j, tk = pending.pop(0)
                got = pr.wait(tk)
                for k in names:
                    assert torch.equal(got[k], want[j][k]), (depth, j, k)
        for j, tk in pending:
            got = pr.wait(tk)
            for k in names:
                assert torch.equal(got[k], want[j][k]), (depth, j, k)
        assert pr.stats()["overflow"] == 0
        with pytest.raises(RuntimeError):
            pr.wait(10_000)
        pr.close()


def test_compositor_without_contributor_count_gives_the_same_images():
    """The compositor instance that does not count contributors (chosen when n_contrib is not requested -- what
    bench.py runs) must produce the images of the counting instance bit for bit."""
    sc = S.make_synthetic_scene(25000, 3, seed=41, link_counts=(1200, 700, 300))
    W, H, E = 160, 112, 3
    vms = np.repeat(S.ring_cameras(1, (0, 0, 0), seed=7), E, axis=0)
    r = SplatRenderer(sc, max_images=E, max_envs=E, max_width=W, max_height=H)
    r.set_link_poses(S.random_link_poses(E, 3, seed=70))
    r.set_cameras(vms, S.fov_intrinsics(W, H), W, H, np.arange(E))
    with_nc = {k: v.clone() for k, v in r.render(("rgb", "depth", "alpha", "rgb_u8", "n_contrib"), background=(0.2, 0.4, 0.6)).items()}
    assert int(with_nc["n_contrib"].max()) > 0
    for _ in range(2):  # direct launch, then the graph replay
        without = r.render(("rgb", "depth", "alpha", "rgb_u8"), background=(0.2, 0.4, 0.6))
        for k in ("rgb", "depth", "alpha", "rgb_u8"):
            assert torch.equal(without[k], with_nc[k]), k
    r.close()


def test_overflow_is_reported_and_regrown():
    sc = S.make_synthetic_scene(30000, 1, seed=17)
    W, H = 128, 128
    vm = S.ring_cameras(1, (0, 0, 0), seed=6)
    r = SplatRenderer(sc, max_images=1, max_width=W, max_height=H, max_isects=2000)
    r.set_cameras(vm, S.fov_intrinsics(W, H), W, H)
    r.render(("rgb",), check=False)
    st = r.stats()
    assert st["overflow"] != 0 and st["n_isects"] > 2000
    out = r.render(("rgb", "depth", "alpha", "rgb_u8", "n_contrib"), check=True)  # grows and re-renders
    assert r.stats()["overflow"] == 0
    ref = oracle.render(sc.as_oracle_dict(), vm, S.fov_intrinsics(W, H)[None], W, H)
    np.testing.assert_array_equal(out["n_contrib"].cpu().numpy(), ref["n_contrib"])
    assert np.abs(out["rgb"].cpu().numpy() - ref["rgb"]).max() <= 1e-3
    r.close()


def test_nerfstudio_post_processing_matches_oracle():
    """GaussianSplat.render contract (nerfstudio_utils.py:123-177): clamped rgb over the splatfacto
    background, expected depth with alpha==0 pixels set to the image max."""
    sc = S.make_synthetic_scene(3000, 3, seed=19)
    W, H = 120, 90
    vm = S.lookat_viewmat((2.5, 0.2, 0.3), (0, 0, 0))[None]  # far enough that corners stay empty
    r = SplatRenderer(sc, max_images=1, max_width=W, max_height=H)
    r.set_cameras(vm, S.fov_intrinsics(W, H, 70.0), W, H)
    out = r.render(("rgb", "depth", "alpha", "rgb_u8"), background=NS_BACKGROUND, nerfstudio_post=True)
    ref = oracle.render(sc.as_oracle_dict(), vm, S.fov_intrinsics(W, H, 70.0)[None], W, H)
    ns = oracle.oracle.nerfstudio_outputs(ref, NS_BACKGROUND)
    assert (ref["alpha"] == 0).sum() > 50
    assert np.abs(out["rgb"].cpu().numpy() - ns["rgb"]).max() <= 1e-3
    assert np.abs(out["depth"].cpu().numpy() - ns["depth"][..., 0]).max() <= 1e-3
    np.testing.assert_array_equal(out["alpha"].cpu().numpy(), ref["alpha"])
    u8 = np.floor(ns["rgb"] * np.float32(255) + np.float32(0.5)).astype(np.uint8)
    assert np.sum(out["rgb_u8"].cpu().numpy() != u8) == 0
    r.close()


def test_full_size_config2_properties(assets):
    """Full BASELINE.json configs[1] size (256 envs x 96x96): a random sample of envs is checked against
    the oracle image-for-image; the rest through size-independent properties."""
    sc = S.make_reference_like_scene("robots-scene-v2", 3, 102, with_tblock=True, assets=assets)
    cen = sc.means[sc.link_id >= 0].mean(0)
    E, W, H = 256, 96, 96
    cam = S.ring_cameras(1, cen, 8)
    poses = S.random_link_poses(E, sc.L, 1000)
    poses[5] = poses[200]  # two envs with identical link poses
    r = SplatRenderer(sc, max_images=E, max_envs=E, max_width=W, max_height=H, max_isects=40_000_000)
    r.set_link_poses(poses)
    r.set_cameras(np.repeat(cam, E, axis=0), S.fov_intrinsics(W, H), W, H, np.arange(E))
    names = ("rgb", "depth", "alpha", "n_contrib")
    a = {k: v.clone() for k, v in r.render(names).items()}
    assert r.stats()["overflow"] == 0
    b = r.render(names)
    for k in names:
        assert torch.equal(a[k], b[k]), k           # idempotent / deterministic despite atomics
        assert torch.equal(a[k][5], a[k][200]), k   # same env state -> same image
    assert not torch.equal(a["rgb"][5], a["rgb"][6])
    # host outputs: composited in 4 image ranges with the device->host copy of each overlapping the next
    # (sas_render, pinned destinations); twice, so that the second call replays the captured graph
    for _ in range(2):
        h = r.render(names + ("rgb_u8",), to_host=True)
        for k in names:
            assert not h[k].is_cuda and torch.equal(h[k], a[k].cpu()), k
    u8 = torch.floor(torch.clamp(a["rgb"], 0, 1) * 255.0 + 0.5).to(torch.uint8).cpu()
    assert torch.equal(h["rgb_u8"], u8)
    pick = [0, 5, 77, 200, 255]
    ref = oracle.render(sc.as_oracle_dict(), np.repeat(cam, len(pick), axis=0), np.repeat(S.fov_intrinsics(W, H)[None], len(pick), 0),
                        W, H, poses[pick], np.arange(len(pick)), keep={"rgb", "depth", "alpha", "n_contrib"})
    for k in names:
        got = a[k][pick].cpu().numpy()
        if k == "n_contrib":
            np.testing.assert_array_equal(got, ref[k])
        else:
            assert np.abs(got - ref[k]).max() <= 1e-3, k
    r.close()
