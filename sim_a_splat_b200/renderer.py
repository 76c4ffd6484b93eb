"""SplatRenderer: host-side object over the C ABI of libsas_b200.so (include/sas.h).

PyTorch is plumbing here (device buffers, pinned host memory, the current CUDA
stream); every pixel is produced by the hand-written sm_100a kernels in csrc/.
There is no CPU or PyTorch fallback: constructing a renderer without a CUDA
device raises.

Reference interfaces this object stands behind (file:line in /root/reference):
  * client.get_render(height,width,wxyz,position)   sim_a_splat/env/splat/splat_env_wrapper.py:148-157
  * handle.wxyz / handle.position per link           sim_a_splat/splat/splat_handler.py:283-288
  * model.get_outputs_for_camera(cameras, obb_box)   sim_a_splat/ns_utils/nerfstudio_utils.py:166-172
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional, Sequence

import numpy as np

from . import _lib as L
from .scene import SplatScene

NS_BACKGROUND = (0.1490, 0.1647, 0.2157)  # splatfacto's eval-time colour for background_color="random" (config.yml:152)


def _f32c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


class SplatRenderer:
    def __init__(self, scene: SplatScene, device: int = 0, max_images: int = 1, max_envs: int = 1,
                 max_width: int = 1920, max_height: int = 1080, max_isects: int = 0):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("sim_a_splat_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self._torch = torch
        self._lib = L.load()
        self.device = int(device)
        self.scene = scene
        self.N, self.L = scene.N, scene.L
        arrs = [_f32c(scene.means), _f32c(scene.quats), _f32c(scene.scales), _f32c(scene.opacities), _f32c(scene.sh),
                np.ascontiguousarray(scene.link_id, dtype=np.int32)]
        desc = L.SceneDesc(scene.N, scene.sh_degree, scene.L, L.SAS_MEM_HOST, *[a.ctypes.data for a in arrs])
        cfg = L.Config(int(max_images), int(max_envs), int(max_width), int(max_height), int(max_isects))
        h = C.c_void_p()
        L.check(self._lib, None, self._lib.sas_create(C.byref(desc), C.byref(cfg), self.device, C.byref(h)))
        self._h = h
        self.max_images, self.max_envs = int(max_images), int(max_envs)
        self.I = self.W = self.H = 0
        # libsas replays a CUDA graph of the step when it is given a real (non-default) stream; when the caller
        # works on torch's default stream the renderer therefore runs on a side stream of its own, ordered
        # against the caller's stream with event waits on both sides.
        self._own_stream = torch.cuda.Stream(device=self.device)
        self._pinned: Dict[str, "torch.Tensor"] = {}

    # -- lifetime ---------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.sas_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self, stream=None) -> C.c_void_p:
        """cudaStream_t for a call.  Explicit stream: used as is.  Otherwise torch's current stream -- unless that is
        the legacy default stream (no graph capture there): then the renderer's own stream, after the caller's work."""
        if stream is None:
            stream = self._torch.cuda.current_stream(self.device)
            if int(stream.cuda_stream) == 0:
                self._own_stream.wait_stream(stream)
                self._handoff = stream
                stream = self._own_stream
        return C.c_void_p(int(stream.cuda_stream))

    def _handback(self) -> None:
        """Make the caller's stream wait for what was enqueued on the renderer's own stream."""
        h = getattr(self, "_handoff", None)
        if h is not None:
            h.wait_stream(self._own_stream)
            self._handoff = None

    # -- per-step inputs --------------------------------------------------------------------
    def set_link_poses(self, poses, stream=None) -> None:
        """poses [E,L,12] (rot row-major | pos): numpy (host) or a CUDA float32 tensor (device)."""
        if self.L == 0:
            return
        torch = self._torch
        if isinstance(poses, torch.Tensor) and poses.is_cuda:
            p = poses.contiguous().float()
            E = p.numel() // (self.L * 12)
            code = self._lib.sas_set_link_poses(self._h, C.c_void_p(p.data_ptr()), E, L.SAS_MEM_DEVICE, self._stream(stream))
            self._keep_poses = p
        else:
            p = _f32c(poses.cpu().numpy() if isinstance(poses, torch.Tensor) else poses).reshape(-1, self.L, 12)
            code = self._lib.sas_set_link_poses(self._h, p.ctypes.data, p.shape[0], L.SAS_MEM_HOST, self._stream(stream))
        L.check(self._lib, self._h, code)
        self._handback()

    def set_pose_model(self, scale: float, Ri, ti, fk_R, fk_t, body_offset=None) -> None:
        """Constants of the reference's link-pose algebra (splat_handler.py:62-83, :158-163, :229-235): after this,
        set_link_poses_from_bodies() composes rot / pos on the device (csrc/sas_pose.cu).  body_offset [L,3] is
        added to each body position (the weld translation for robot bodies, zero for free bodies)."""
        fkR = np.ascontiguousarray(fk_R, dtype=np.float64).reshape(self.L, 9)
        fkt = np.ascontiguousarray(fk_t, dtype=np.float64).reshape(self.L, 3)
        off = np.zeros((self.L, 3)) if body_offset is None else np.ascontiguousarray(body_offset, dtype=np.float64).reshape(self.L, 3)
        m = L.PoseModel()
        m.n_links, m.scale = self.L, float(scale)
        m.Ri = (C.c_double * 9)(*np.asarray(Ri, dtype=np.float64).reshape(9))
        m.ti = (C.c_double * 3)(*np.asarray(ti, dtype=np.float64).reshape(3))
        m.fk_R, m.fk_t, m.body_offset = fkR.ctypes.data, fkt.ctypes.data, off.ctypes.data
        L.check(self._lib, self._h, self._lib.sas_set_pose_model(self._h, C.byref(m)))

    def set_link_poses_from_bodies(self, body_wxyz_xyz, stream=None) -> None:
        """body_wxyz_xyz [E,L,7] fp64: the draw message's (quaternion wxyz, position) of the L robot bodies that
        drive the link groups (manipulator_env.py:377-390), numpy (host) or a CUDA float64 tensor."""
        torch = self._torch
        if isinstance(body_wxyz_xyz, torch.Tensor) and body_wxyz_xyz.is_cuda:
            b = body_wxyz_xyz.contiguous().double()
            code = self._lib.sas_set_link_poses_from_bodies(self._h, C.c_void_p(b.data_ptr()), b.numel() // (self.L * 7), L.SAS_MEM_DEVICE,
                                                            self._stream(stream))
            self._keep_bodies = b
        else:
            b = np.ascontiguousarray(body_wxyz_xyz, dtype=np.float64).reshape(-1, self.L, 7)
            code = self._lib.sas_set_link_poses_from_bodies(self._h, b.ctypes.data, b.shape[0], L.SAS_MEM_HOST, self._stream(stream))
        L.check(self._lib, self._h, code)
        self._handback()

    def unproject_depth(self, depth, K, max_depth: float = 1e10, stream=None):
        """Pinhole unprojection of nerfstudio_utils.py:441-452 on the device: depth [I,H,W] (CUDA float32 tensor) ->
        (xyz [I,H,W,3] camera-frame points, valid [I,H,W] bool = depth < max_depth)."""
        torch = self._torch
        d = depth.contiguous().float()
        if d.dim() == 2:
            d = d[None]
        I, H, W = d.shape
        xyz = torch.empty((I, H, W, 3), dtype=torch.float32, device=d.device)
        valid = torch.empty((I, H, W), dtype=torch.uint8, device=d.device)
        Kf = _f32c(K).reshape(9)
        L.check(self._lib, self._h, self._lib.sas_unproject_depth(self._h, C.c_void_p(d.data_ptr()), I, W, H, Kf.ctypes.data, float(max_depth),
                                                                  C.c_void_p(xyz.data_ptr()), C.c_void_p(valid.data_ptr()), L.SAS_MEM_DEVICE,
                                                                  self._stream(stream)))
        self._handback()
        return xyz, valid.bool()

    def set_cameras(self, viewmats, Ks, width: int, height: int, env_of_image: Optional[Sequence[int]] = None, stream=None) -> None:
        """viewmats [I,4,4] world->camera (OpenCV axes), Ks [I,3,3] or [3,3]; host arrays."""
        vm = _f32c(viewmats).reshape(-1, 16)
        I = vm.shape[0]
        K = _f32c(Ks).reshape(-1, 9)
        if K.shape[0] == 1 and I > 1:
            K = np.ascontiguousarray(np.repeat(K, I, axis=0))
        env = None
        if env_of_image is not None:
            env = np.ascontiguousarray(env_of_image, dtype=np.int32)
            assert env.shape == (I,)
        code = self._lib.sas_set_cameras(self._h, vm.ctypes.data, K.ctypes.data, env.ctypes.data if env is not None else None,
                                         I, int(width), int(height), self._stream(stream))
        L.check(self._lib, self._h, code)
        self._handback()
        self.I, self.W, self.H = I, int(width), int(height)

    # -- render -----------------------------------------------------------------------------
    def _alloc(self, name: str, to_host: bool, slot: int = 0):
        torch = self._torch
        I, H, W = self.I, self.H, self.W
        shape, dt = {"rgb": ((I, H, W, 3), torch.float32), "depth": ((I, H, W), torch.float32),
                     "alpha": ((I, H, W), torch.float32), "rgb_u8": ((I, H, W, 3), torch.uint8),
                     "n_contrib": ((I, H, W), torch.int32)}[name]
        if to_host:
            t = self._pinned.get((name, slot))
            if t is None or tuple(t.shape) != shape:
                t = torch.empty(shape, dtype=dt, pin_memory=True)
                self._pinned[(name, slot)] = t
            return t
        return torch.empty(shape, dtype=dt, device=f"cuda:{self.device}")

    def render(self, outputs: Iterable[str] = ("rgb", "depth", "alpha"), background=(0.0, 0.0, 0.0),
               nerfstudio_post: bool = False, to_host: bool = False, out: Optional[dict] = None, stream=None,
               check: bool = True) -> dict:
        """Render every image set by set_cameras.  Returns {name: torch tensor}: CUDA tensors, or (to_host=True)
        pinned host tensors filled by a D2H copy inside the call (which then synchronises the stream).
        check=True reads the device counters afterwards (a sync) and transparently grows the intersection
        buffers and re-renders if they overflowed."""
        if self.I == 0:
            raise RuntimeError("set_cameras has not been called")
        names = list(outputs)
        bufs = dict(out) if out else {}
        for n in names:
            if n not in bufs:
                bufs[n] = self._alloc(n, to_host)
        o = L.Outputs()
        o.mem = L.SAS_MEM_HOST if to_host else L.SAS_MEM_DEVICE
        o.nerfstudio_post = 1 if nerfstudio_post else 0
        o.background = (C.c_float * 3)(*[float(b) for b in background])
        for n in ("rgb", "depth", "alpha", "rgb_u8", "n_contrib"):
            setattr(o, n, C.c_void_p(bufs[n].data_ptr()) if n in bufs else None)
        st = self._stream(stream)
        L.check(self._lib, self._h, self._lib.sas_render(self._h, C.byref(o), st))
        if check:
            s = self.stats()
            if s["overflow"]:
                L.check(self._lib, self._h, self._lib.sas_reserve_isects(self._h, int(s["n_isects"] * 1.25) + 1024))
                L.check(self._lib, self._h, self._lib.sas_render(self._h, C.byref(o), st))
                if self.stats()["overflow"]:
                    raise L.SasError(-4, "intersection buffers overflowed twice")
        self._handback()
        return {n: bufs[n] for n in names}

    def render_async(self, outputs: Iterable[str] = ("rgb_u8", "depth"), background=(0.0, 0.0, 0.0),
                     nerfstudio_post: bool = False, stream=None) -> int:
        """Issue a render into pinned host buffers WITHOUT waiting for it (sas_outputs.mem = SAS_MEM_HOST_ASYNC): the
        D2H copies run on the library's copy stream out of one of two alternating staging sets, so the transfer of
        step k overlaps the render of step k+1.  Returns a ticket for wait().  At most two renders may be outstanding
        (issue k+1, then wait(k)); the tensors wait(k) returned are reused by render k+2.  No overflow check: size
        max_isects beforehand (stats()['overflow'] after a render tells)."""
        if self.I == 0:
            raise RuntimeError("set_cameras has not been called")
        seq = getattr(self, "_async_seq", 0)
        names = list(outputs)
        bufs = {n: self._alloc(n, True, slot=seq & 1) for n in names}
        o = L.Outputs()
        o.mem = L.SAS_MEM_HOST_ASYNC
        o.nerfstudio_post = 1 if nerfstudio_post else 0
        o.background = (C.c_float * 3)(*[float(b) for b in background])
        for n in ("rgb", "depth", "alpha", "rgb_u8", "n_contrib"):
            setattr(o, n, C.c_void_p(bufs[n].data_ptr()) if n in bufs else None)
        L.check(self._lib, self._h, self._lib.sas_render(self._h, C.byref(o), self._stream(stream)))
        self._handback()
        if not hasattr(self, "_async_bufs"):
            self._async_bufs = [None, None]
        self._async_bufs[seq & 1] = bufs
        self._async_seq = seq + 1
        return seq

    def wait(self, ticket: Optional[int] = None) -> dict:
        """Block until the host buffers of render_async ticket (default: the latest) are complete; returns them."""
        seq = getattr(self, "_async_seq", 0)
        if ticket is None:
            ticket = seq - 1
        age = seq - 1 - ticket
        if seq == 0 or age < 0:
            raise RuntimeError("wait(): no such render_async ticket")
        if age <= 1:  # older renders have completed by construction (their staging set was reused since)
            L.check(self._lib, self._h, self._lib.sas_wait_outputs(self._h, int(age)))
        else:
            raise RuntimeError("wait(): the buffers of this ticket have been reused by a later render_async")
        return self._async_bufs[ticket & 1]

    def reserve_isects(self, n: int) -> None:
        L.check(self._lib, self._h, self._lib.sas_reserve_isects(self._h, int(n)))

    def stats(self) -> dict:
        s = L.Stats()
        L.check(self._lib, self._h, self._lib.sas_get_stats(self._h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in L.Stats._fields_}

    STAGES = ("project", "scan", "emit", "sort", "composite", "post")

    def set_stage_timing(self, enable: bool) -> None:
        L.check(self._lib, self._h, self._lib.sas_set_stage_timing(self._h, 1 if enable else 0))

    def stage_ms(self) -> dict:
        ms = (C.c_float * len(self.STAGES))()
        L.check(self._lib, self._h, self._lib.sas_get_stage_ms(self._h, ms))
        return dict(zip(self.STAGES, [float(v) for v in ms]))

    def last_launch_count(self) -> int:
        return int(self._lib.sas_last_launch_count(self._h))

    def link_poses(self, n_envs: int) -> np.ndarray:
        """[n_envs, L, 12] fp32 link poses currently in effect on the device (parity tests)."""
        out = np.empty((n_envs, self.L, 12), dtype=np.float32)
        L.check(self._lib, self._h, self._lib.sas_debug_fetch(self._h, L.SAS_DBG_LINK_POSES, out.ctypes.data, out.nbytes))
        return out

    # -- parity-test access to intermediates ------------------------------------------------
    def debug_intermediates(self) -> dict:
        """Intermediates of the last render in the oracle's per-(image, gaussian) layout (numpy, host).
        The library keeps background records once per view and link records once per image; they are
        re-assembled here."""
        s = self.stats()
        I, N, C, nbg = self.I, self.N, s["n_views"], s["n_background"]
        ndy = N - nbg
        tw, th = s["tile_w"], s["tile_h"]
        T = tw * th
        M = min(s["n_isects"], 2**31 - 16)
        n_rec = C * nbg + I * ndy
        rec_raw = np.empty((n_rec, 12), dtype=np.float32)
        bins_raw = np.empty((n_rec, 2), dtype=np.uint32)
        radii_raw = np.empty((n_rec, 2), dtype=np.uint32)
        offs = np.empty(I * T + 1, dtype=np.int32)
        gids = np.empty(M, dtype=np.int32)
        voi = np.empty(I, dtype=np.int32)
        for what, arr in ((L.SAS_DBG_RECORDS, rec_raw), (L.SAS_DBG_BINS, bins_raw), (L.SAS_DBG_RADII, radii_raw), (L.SAS_DBG_TILE_OFFSETS, offs),
                          (L.SAS_DBG_SORTED_GIDS, gids), (L.SAS_DBG_VIEW_OF_IMAGE, voi)):
            L.check(self._lib, self._h, self._lib.sas_debug_fetch(self._h, what, arr.ctypes.data, arr.nbytes))

        def per_image(raw):
            bg = raw[: C * nbg].reshape(C, nbg, raw.shape[1])[voi]
            dy = raw[C * nbg:].reshape(I, ndy, raw.shape[1])
            return np.concatenate([bg, dy], axis=1)

        rec, bins, radii = per_image(rec_raw), per_image(bins_raw), per_image(radii_raw).astype(np.int32)
        rect = np.stack([bins[..., 0] & 0xff, (bins[..., 0] >> 8) & 0xff, (bins[..., 0] >> 16) & 0xff, bins[..., 0] >> 24], axis=-1).astype(np.int32)
        vis = bins[..., 0] != 0
        tiles = ((rect[..., 2] - rect[..., 0]) * (rect[..., 3] - rect[..., 1])).astype(np.int32)
        z = lambda a: np.where(vis[..., None] if a.ndim == 3 else vis, a, 0).astype(a.dtype)  # noqa: E731
        two = np.float32(2.0)
        conics = np.stack([rec[..., 2] * two, rec[..., 3], rec[..., 4] * two], axis=-1)
        tile_bits = int(T).bit_length()
        # rebuild gsplat's 64-bit keys / flattened ids from the bucket structure
        counts = np.diff(offs.astype(np.int64))
        bucket = np.repeat(np.arange(I * T, dtype=np.int64), counts)[:M]
        img, tile = bucket // T, bucket % T
        flat = (img * N + gids.astype(np.int64))
        depth_bits = bins.reshape(-1, 2)[flat, 1].astype(np.int64)
        isect_ids = (img << (32 + tile_bits)) | (tile << 32) | depth_bits
        return dict(
            radii=z(radii), extents=z(rec[..., 6:8]), means2d=z(rec[..., 0:2]), depths=z(rec[..., 11]), conics=z(conics),
            colors=z(rec[..., 8:11]), opacities=z(rec[..., 5]), tiles_per_gauss=z(tiles), tile_rect=z(rect), visible=vis,
            isect_ids=isect_ids, flatten_ids=flat.astype(np.int32), isect_offsets=offs[:-1].reshape(I, th, tw).copy(),
            M=M, tile_bits=tile_bits, stats=s, view_of_image=voi,
        )


class PipelinedSplatRenderer:
    """`depth` SplatRenderers (each its own libsas handle, work buffers and CUDA stream) used round-robin, so that
    consecutive steps overlap on the device as well as on the copy engines: while step k composites (its tail is the
    latency of the heaviest tile's blend chain), step k+1 already projects, bins and sorts on the
    SMs the compositor no longer fills, and step k-1's observations travel to the host.  Every step still does all of
    its own work; the images are bit-identical to SplatRenderer's (tests/test_gpu_parity.py).

    Use: set_cameras once; per step set_link_poses(...) (or set_link_poses_from_bodies) then render_async() -> ticket;
    wait(ticket) returns that step's pinned host tensors.  Up to 2*depth steps may be outstanding (every handle
    double-buffers its staging and host buffers): issuing step k and then waiting for step k-depth keeps every handle's
    next render queued behind its current one.  The tensors of step k are reused by step k + 2*depth.

    The reference has no counterpart (it blocks on one browser round trip per camera,
    /root/reference/sim_a_splat/env/splat/splat_env_wrapper.py:147-158); this is the step_async / step_wait split of
    vectorised envs applied to the observation render.  The scene is held once per handle (27-70 MB for the
    reference's scenes)."""

    def __init__(self, scene: SplatScene, depth: int = 2, device: int = 0, **kw):
        import torch

        if depth < 1:
            raise ValueError("depth must be >= 1")
        self._torch = torch
        self.device = int(device)
        self.renderers = [SplatRenderer(scene, device=device, **kw) for _ in range(depth)]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(depth)]
        self.depth, self.L, self.N = depth, scene.L, scene.N
        self._seq = 0          # steps issued so far; step k runs on renderer k % depth
        self._sub = {}         # ticket -> sub-renderer ticket

    def _cur(self):
        k = self._seq % self.depth
        return self.renderers[k], self.streams[k]

    def _after_caller(self, s) -> None:
        s.wait_stream(self._torch.cuda.current_stream(self.device))  # inputs produced on the caller's stream

    def set_cameras(self, viewmats, Ks, width: int, height: int, env_of_image=None, current_only: bool = False) -> None:
        """Cameras of every handle (fixed cameras: call once), or with current_only=True of the NEXT render_async only
        (cameras that move with the step: call before each render_async)."""
        for r, s in [self._cur()] if current_only else zip(self.renderers, self.streams):
            r.set_cameras(viewmats, Ks, width, height, env_of_image, stream=s)

    def set_pose_model(self, **kw) -> None:
        for r in self.renderers:
            r.set_pose_model(**kw)

    def set_link_poses(self, poses) -> None:
        """Link poses of the NEXT render_async."""
        r, s = self._cur()
        self._after_caller(s)
        r.set_link_poses(poses, stream=s)

    def set_link_poses_from_bodies(self, body_wxyz_xyz) -> None:
        r, s = self._cur()
        self._after_caller(s)
        r.set_link_poses_from_bodies(body_wxyz_xyz, stream=s)

    def render_async(self, outputs: Iterable[str] = ("rgb_u8", "depth"), background=(0.0, 0.0, 0.0),
                     nerfstudio_post: bool = False) -> int:
        r, s = self._cur()
        self._after_caller(s)
        self._sub[self._seq] = r.render_async(outputs, background=background, nerfstudio_post=nerfstudio_post, stream=s)
        self._sub.pop(self._seq - 2 * self.depth, None)
        self._seq += 1
        return self._seq - 1

    def render(self, outputs: Iterable[str] = ("rgb", "depth", "alpha"), background=(0.0, 0.0, 0.0),
               nerfstudio_post: bool = False) -> dict:
        """Device-resident outputs of the next step, enqueued on that step's own stream; returns at once.  The CUDA
        tensors belong to the handle that drew them (valid until its next render, `depth` steps later) and are ordered
        on its stream: call join() before consuming them on the caller's stream."""
        r, s = self._cur()
        self._after_caller(s)
        k = self._seq % self.depth
        names = tuple(outputs)
        if not hasattr(self, "_dev_out"):
            self._dev_out = {}
        bufs = self._dev_out.get((k, names))
        if bufs is None or any(tuple(b.shape[:3]) != (r.I, r.H, r.W) for b in bufs.values()):
            bufs = self._dev_out[(k, names)] = {n: r._alloc(n, False) for n in names}
        r.render(names, background=background, nerfstudio_post=nerfstudio_post, out=bufs, stream=s, check=False)
        self._seq += 1
        return bufs

    def join(self) -> None:
        """Order the caller's current stream after every step issued so far."""
        cur = self._torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)

    def wait(self, ticket: Optional[int] = None) -> dict:
        if ticket is None:
            ticket = self._seq - 1
        if ticket not in self._sub:
            raise RuntimeError("wait(): no such render_async ticket (never issued, or its buffers have been reused)")
        return self.renderers[ticket % self.depth].wait(self._sub[ticket])

    def stats(self) -> dict:
        """Counters of the most recently issued step's renderer."""
        return self.renderers[(self._seq - 1) % self.depth].stats()

    def last_launch_count(self) -> int:
        return self.renderers[(self._seq - 1) % self.depth].last_launch_count()

    def close(self) -> None:
        for r in self.renderers:
            r.close()
