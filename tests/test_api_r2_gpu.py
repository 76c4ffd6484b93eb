"""Round-2 boundary features, through the C ABI on the GPU: one-copy packed host outputs and the fp16 depth output,
chunked rendering of more images than one library call takes, overflow that is reported instead of silently
truncating asynchronous renders, env-index validation, the input lifetime of pose tensors handed to a renderer that
works on its own stream, separate capacities for sorted bucket entries and merged ids."""
import ctypes as C

import numpy as np
import pytest
import torch

from sim_a_splat_b200 import _lib as L
from sim_a_splat_b200 import scene as S
from sim_a_splat_b200.renderer import PipelinedSplatRenderer, SplatRenderer

pytestmark = pytest.mark.gpu


def _scene():
    return S.make_synthetic_scene(9000, 3, seed=77, link_counts=(400, 250, 150))


def _setup(r, sc, E, W, H, n_cams=2, seed=5):
    cams = S.ring_cameras(n_cams, (0, 0, 0), seed=seed)
    vms = np.tile(cams, (E, 1, 1))
    env = np.repeat(np.arange(E, dtype=np.int32), n_cams)
    poses = S.random_link_poses(E, sc.L, seed=900)
    r.set_link_poses(poses)
    r.set_cameras(vms, S.fov_intrinsics(W, H), W, H, env)
    return vms, env, poses


def test_packed_host_outputs_one_block_and_fp16_depth():
    sc = _scene()
    W, H, E = 80, 64, 3
    r = SplatRenderer(sc, max_images=2 * E, max_envs=E, max_width=W, max_height=H)
    _setup(r, sc, E, W, H)
    names = ("rgb_u8", "depth_f16", "depth", "alpha", "rgb", "n_contrib")
    dev = {k: v.cpu() for k, v in r.render(names).items()}
    host = r.render(names, to_host=True)
    blk = r._alloc_packed(list(names))["_block"]
    lo, hi = blk.data_ptr(), blk.data_ptr() + blk.numel()
    for k in names:
        assert host[k].is_pinned() and lo <= host[k].data_ptr() < hi, k       # views of ONE pinned block
        assert host[k].data_ptr() % 256 == lo % 256, k                         # the library's 256-byte aligned layout
        assert torch.equal(host[k], dev[k]), k
    # fp16 depth = round-to-nearest-even of the f32 expected depth (what torch's .half() does)
    assert host["depth_f16"].dtype == torch.float16
    assert torch.equal(host["depth_f16"].view(torch.int16), dev["depth"].half().view(torch.int16))
    # a subset, and host buffers that are NOT one block (per-output copies) give the same bytes
    sub = r.render(("rgb_u8", "depth"), to_host=True)
    assert torch.equal(sub["rgb_u8"], dev["rgb_u8"]) and torch.equal(sub["depth"], dev["depth"])
    loose = {"rgb_u8": torch.empty_like(dev["rgb_u8"]).pin_memory(), "depth": torch.empty_like(dev["depth"]).pin_memory()}
    got = r.render(("rgb_u8", "depth"), to_host=True, out=loose)
    assert torch.equal(got["rgb_u8"], dev["rgb_u8"]) and torch.equal(got["depth"], dev["depth"])
    # nerfstudio post-processing keeps the fp16 copy in step with the f32 depth it rewrites where alpha == 0
    ns = r.render(("depth", "depth_f16"), nerfstudio_post=True, background=(0.1, 0.2, 0.3))
    assert torch.equal(ns["depth_f16"].view(torch.int16), ns["depth"].half().view(torch.int16))
    r.close()


def test_async_host_outputs_match_blocking():
    sc = _scene()
    W, H, E = 64, 48, 4
    r = SplatRenderer(sc, max_images=2 * E, max_envs=E, max_width=W, max_height=H)
    _setup(r, sc, E, W, H)
    ref = {k: v.cpu() for k, v in r.render(("rgb_u8", "depth_f16")).items()}
    tickets = [r.render_async(("rgb_u8", "depth_f16")) for _ in range(2)]
    for t in tickets:
        got = r.wait(t)
        assert torch.equal(got["rgb_u8"], ref["rgb_u8"]) and torch.equal(got["depth_f16"].view(torch.int16), ref["depth_f16"].view(torch.int16))
    r.close()


def test_chunked_render_reproduces_one_call_bit_for_bit():
    sc = _scene()
    W, H, E = 96, 80, 5
    names = ("rgb", "depth", "alpha", "rgb_u8", "n_contrib")
    full = SplatRenderer(sc, max_images=2 * E, max_envs=E, max_width=W, max_height=H)
    _setup(full, sc, E, W, H)
    a = {k: v.clone() for k, v in full.render(names).items()}
    full.close()
    ch = SplatRenderer(sc, max_images=3, max_envs=E, max_width=W, max_height=H)   # 10 images in calls of 3, 3, 3, 1
    _setup(ch, sc, E, W, H)
    b = ch.render(names)
    for k in names:
        assert torch.equal(a[k], b[k]), k
    h = ch.render(("rgb_u8", "depth"), to_host=True)
    assert torch.equal(h["rgb_u8"], a["rgb_u8"].cpu()) and torch.equal(h["depth"], a["depth"].cpu())
    with pytest.raises(RuntimeError):
        ch.render_async(("rgb_u8",))
    ch.close()


def test_overflow_is_reported_not_silent():
    sc = _scene()
    W, H, E = 96, 96, 4
    r = SplatRenderer(sc, max_images=2 * E, max_envs=E, max_width=W, max_height=H, max_isects=2048, max_isects_src=1024)
    _setup(r, sc, E, W, H)
    st0 = r.stats()
    assert st0["cap_isects"] == 2048 and st0["cap_isects_src"] == 1024
    # (a) unchecked device render: nothing raises at once, poll_overflow() tells and names the capacities that suffice
    r.render(("rgb",), check=False)
    need = r.poll_overflow()
    assert need is not None and need["n_isects"] > 2048 and need["n_isects_src"] > 1024
    assert r.poll_overflow() is None                       # reported once
    # (b) blocking host render: SAS_ECAPACITY from the library itself
    o = r._outputs_struct({"rgb_u8": r._alloc("rgb_u8", True)}, L.SAS_MEM_HOST, False, (0, 0, 0))
    code = r._lib.sas_render(r._h, C.byref(o), r._stream(None))
    r._handback()
    assert code == -4 and b"INCOMPLETE" in r._lib.sas_last_error(r._h)
    # (c) asynchronous host render whose buffers are too small: wait() raises
    r._sized = True                                        # pretend an earlier step had sized the buffers
    t = r.render_async(("rgb_u8",))
    with pytest.raises(L.SasError) as ei:
        r.wait(t)
    assert ei.value.code == -4
    # (d) the normal path: the first asynchronous render sizes the buffers with one checked render, separately
    r._sized = False
    got = r.wait(r.render_async(("rgb_u8",)))
    st = r.stats()
    assert st["overflow"] == 0 and st["cap_isects"] >= st["n_isects"] and st["cap_isects_src"] >= st["n_isects_src"]
    assert st["cap_isects_src"] < st["cap_isects"]         # shared views: far fewer sorted entries than merged ids
    ref = r.render(("rgb_u8",))["rgb_u8"].cpu()
    assert torch.equal(got["rgb_u8"], ref)
    r.close()


def test_env_index_beyond_uploaded_poses_is_rejected():
    sc = _scene()
    W, H = 48, 48
    r = SplatRenderer(sc, max_images=4, max_envs=4, max_width=W, max_height=H)
    cams = S.ring_cameras(1, (0, 0, 0), seed=3)
    r.set_link_poses(S.random_link_poses(2, sc.L, seed=1))                  # poses of envs 0 and 1 only
    r.set_cameras(np.repeat(cams, 3, axis=0), S.fov_intrinsics(W, H), W, H, np.array([0, 1, 3], np.int32))
    with pytest.raises(L.SasError) as ei:
        r.render(("rgb",))
    assert ei.value.code == -5
    r.set_link_poses(S.random_link_poses(4, sc.L, seed=1))
    r.render(("rgb",))
    r.close()


def test_pose_tensor_may_be_overwritten_right_after_the_call():
    """a renderer that works on its own stream reads the caller's pose tensor asynchronously; the Python
    layer takes a private copy on the caller's stream, so an in-place update for the next step cannot race it."""
    sc = _scene()
    W, H, E = 64, 64, 8
    pr = PipelinedSplatRenderer(sc, depth=2, max_images=E, max_envs=E, max_width=W, max_height=H)
    cams = S.ring_cameras(1, (0, 0, 0), seed=9)
    pr.set_cameras(np.repeat(cams, E, axis=0), S.fov_intrinsics(W, H), W, H, np.arange(E))
    pa, pb = (torch.from_numpy(S.random_link_poses(E, sc.L, seed=s)).cuda() for s in (11, 12))
    ref = SplatRenderer(sc, max_images=E, max_envs=E, max_width=W, max_height=H)
    ref.set_cameras(np.repeat(cams, E, axis=0), S.fov_intrinsics(W, H), W, H, np.arange(E))
    want = []
    for p in (pa, pb):
        ref.set_link_poses(p)
        want.append(ref.render(("rgb",))["rgb"].clone())
    buf = pa.clone()
    big = torch.empty(64 << 20, device="cuda")
    outs = []
    for k in range(6):
        src = pa if k % 2 == 0 else pb
        buf.copy_(src)
        big.normal_()                      # keeps the caller's stream busy so the handle's stream runs behind it
        pr.set_link_poses(buf)
        buf.zero_()                        # the caller reuses its tensor immediately
        outs.append(pr.render(("rgb",))["rgb"])
        pr.join()
        assert torch.equal(outs[-1], want[k % 2]), k
    pr.close(); ref.close()


def test_radius_rule_cuda_1_5_fewer_intersections_same_images():
    """SURVEY 8c pinned rules: gsplat 1.5's CUDA projection shrinks the radius extend with the opacity and culls
    opacity < 1/255; as an optional mode (sas_config.radius_rule = SAS_RADIUS_CUDA_1_5).  Integer stages bit-exact
    against the oracle run with the same rule; images and contributor counts identical to the default rule."""
    import oracle
    from parity_utils import INT_KEYS

    sc = S.make_synthetic_scene(20000, 3, seed=41, link_counts=(700, 400))
    rng = np.random.default_rng(4)
    sc.opacities[rng.choice(sc.N, 3000, replace=False)] = rng.uniform(0.001, 0.05, 3000).astype(np.float32)  # incl. below 1/255
    W, H, E = 200, 144, 2
    cams = S.ring_cameras(2, (0, 0, 0), seed=6)
    vms = np.tile(cams, (E, 1, 1))
    env = np.repeat(np.arange(E, dtype=np.int32), 2)
    K = S.fov_intrinsics(W, H)
    poses = S.random_link_poses(E, sc.L, seed=77)
    Ks = np.repeat(K[None], len(vms), 0)
    ref0 = oracle.render(sc.as_oracle_dict(), vms, Ks, W, H, poses, env)
    ref1 = oracle.render(sc.as_oracle_dict(), vms, Ks, W, H, poses, env, radius_rule="cuda_1_5")
    assert ref1["M"] < ref0["M"]
    for k in ("rgb", "depth", "alpha", "n_contrib"):
        np.testing.assert_array_equal(ref0[k], ref1[k])                # the rule cannot change a pixel
    r = SplatRenderer(sc, max_images=len(vms), max_envs=E, max_width=W, max_height=H, radius_rule="cuda_1_5")
    r.set_link_poses(poses)
    r.set_cameras(vms, K, W, H, env)
    out = r.render(("rgb", "depth", "alpha", "n_contrib"))
    dbg = r.debug_intermediates()
    assert dbg["M"] == ref1["M"]
    for k in INT_KEYS:
        np.testing.assert_array_equal(np.asarray(dbg[k]), np.asarray(ref1[k]), err_msg=k)
    for k in ("rgb", "depth", "alpha", "n_contrib"):
        np.testing.assert_array_equal(out[k].cpu().numpy(), ref1[k], err_msg=k)
    r.close()


_COMPACT_PROBE = r"""
import hashlib, json, sys
import numpy as np
from sim_a_splat_b200 import scene as S
from sim_a_splat_b200.renderer import SplatRenderer
sc = S.make_synthetic_scene(30000, 3, seed=23, link_counts=(900, 500, 200))
W, H, E = 96, 96, 12
vms = np.repeat(S.ring_cameras(1, (0, 0, 0), seed=4), E, axis=0)
r = SplatRenderer(sc, max_images=E, max_envs=E, max_width=W, max_height=H)
r.set_link_poses(S.random_link_poses(E, 3, seed=31))
r.set_cameras(vms, S.fov_intrinsics(W, H), W, H, np.arange(E))
out = r.render(("rgb", "depth", "alpha", "n_contrib"), background=(0.3, 0.1, 0.7))
h = hashlib.sha256()
for k in ("rgb", "depth", "alpha", "n_contrib"):
    h.update(out[k].cpu().numpy().tobytes())
print(json.dumps({"sha": h.hexdigest(), "n_staged": r.stats()["n_staged"], "n_isects": r.stats()["n_isects"]}))
"""


def test_compacted_compositor_lists_same_bits_fewer_staged_records():
    """SAS_VARIANT bits 12 and 13 (default on) and 14 (ablation): background records that reach none of the groups a tile's link Gaussians can reach
    are left out of that tile's compositor list.  Against a process with the bit off: fewer staged records, the same
    images / depth / alpha / contributor counts bit for bit; and the compacted render matches the oracle."""
    import json, os, subprocess, sys
    import oracle
    from parity_utils import assert_parity, compare, run_both

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for name, var in (("on", "0x72cf"), ("vrec", "0x32cf"), ("ids", "0x12cf"), ("off", "0x2cf"), ("gate", "0x32cf")):
        env = dict(os.environ, SAS_VARIANT=var, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
        if name == "gate":   # the per-tile ablation gate: tiles above 8 reachable groups stay uncompacted, but still name their
            env["SAS_COMPACT_MAX"] = "8"   # background records by position (a fuzz sweep caught this combination reading ids as positions)
        p = subprocess.run([sys.executable, "-c", _COMPACT_PROBE], env=env, capture_output=True, text=True, timeout=600, cwd=root)
        assert p.returncode == 0, p.stderr[-2000:]
        res[name] = json.loads(p.stdout.strip().splitlines()[-1])
    # "on": bits 12 + 13 + 14 (the ablation with a list of its own for each of the two CTAs of a tile with more than 16 groups);
    # "vrec": bits 12 + 13 (compacted lists naming the view buckets' contiguous records, exact background reach masks);
    # "ids": bit 12 alone (compacted id lists, culling-box masks); "off": none of them
    for k in ("on", "vrec", "ids", "gate"):
        assert res[k]["sha"] == res["off"]["sha"], k
    assert res["vrec"]["n_staged"] <= res["gate"]["n_staged"] <= res["off"]["n_staged"], res
    assert res["on"]["n_isects"] == res["off"]["n_isects"] == res["ids"]["n_isects"] == res["vrec"]["n_isects"]
    assert res["ids"]["n_staged"] < 0.95 * res["off"]["n_staged"], res
    assert res["vrec"]["n_staged"] <= res["ids"]["n_staged"], res      # exact masks can only drop more
    assert res["on"]["n_staged"] <= res["vrec"]["n_staged"], res       # a part's list is a subset of the tile's
    # the same shape in this process (whatever SAS_VARIANT it runs under) against the oracle
    sc = S.make_synthetic_scene(30000, 3, seed=23, link_counts=(900, 500, 200))
    W, H, E = 96, 96, 12
    vms = np.repeat(S.ring_cameras(1, (0, 0, 0), seed=4), E, axis=0)
    gpu, ref = run_both(sc, vms, S.fov_intrinsics(W, H)[None], W, H, S.random_link_poses(E, 3, seed=31), np.arange(E), bg=(0.3, 0.1, 0.7))
    assert_parity(compare(gpu, ref))
