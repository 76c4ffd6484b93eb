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
    """One libsas handle.  `max_images` is the capacity of ONE library call; set_cameras() accepts more images than
    that and render() then walks them in chunks (cameras are the most significant field of gsplat's sort key, so the
    chunks reproduce the one-call images bit for bit).  `max_isects` / `max_isects_src` size the intersection buffers
    (merged ids / sorted bucket entries); 0 = start small and let the first (checked) render grow them."""

    def __init__(self, scene: SplatScene, device: int = 0, max_images: int = 1, max_envs: int = 1,
                 max_width: int = 1920, max_height: int = 1080, max_isects: int = 0, max_isects_src: int = 0,
                 radius_rule: str = "torch"):
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
        cfg = L.Config(int(max_images), int(max_envs), int(max_width), int(max_height), int(max_isects), int(max_isects_src),
                       {"torch": 0, "cuda_1_5": 1}[radius_rule], 0)
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
        self._sized = False        # True once a checked render has sized the intersection buffers for the current cameras
        self._cams = None          # host copy of the camera set (chunked rendering re-uploads slices of it)
        self._pose_stage = None    # private device copy of the caller's pose tensor (foreign-stream calls)
        self._pose_read_ev = None

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
    def _private_copy(self, t, stream, slot: str):
        """The library reads pose tensors asynchronously on `stream` (include/sas.h, input lifetime).  When that is not
        the caller's current stream nothing would order the caller's next in-place update after that read, so the
        tensor is first copied, ON THE CALLER'S STREAM, into a buffer this renderer owns; `stream` then waits for the
        copy.  The caller may overwrite its tensor as soon as this returns.  (Calls on the caller's own stream, or on
        the renderer's side stream with its hand-back, are ordered already.)"""
        torch = self._torch
        if stream is None:
            return t
        cur = torch.cuda.current_stream(self.device)
        if int(stream.cuda_stream) == int(cur.cuda_stream):
            return t
        buf = getattr(self, slot, None)
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            buf = torch.empty_like(t)
            setattr(self, slot, buf)
        elif self._pose_read_ev is not None:
            cur.wait_event(self._pose_read_ev)   # the previous read of the private buffer (issued steps ago) has run
        buf.copy_(t)                             # on the caller's stream
        stream.wait_stream(cur)
        return buf

    def _after_private_read(self, stream) -> None:
        if stream is not None and int(stream.cuda_stream) != int(self._torch.cuda.current_stream(self.device).cuda_stream):
            self._pose_read_ev = stream.record_event()

    def set_link_poses(self, poses, stream=None) -> None:
        """poses [E,L,12] (rot row-major | pos): numpy (host) or a CUDA float32 tensor (device)."""
        if self.L == 0:
            return
        torch = self._torch
        if isinstance(poses, torch.Tensor) and poses.is_cuda:
            p = self._private_copy(poses.contiguous().float(), stream, "_pose_stage")
            E = p.numel() // (self.L * 12)
            code = self._lib.sas_set_link_poses(self._h, C.c_void_p(p.data_ptr()), E, L.SAS_MEM_DEVICE, self._stream(stream))
            self._keep_poses = p
            self._after_private_read(stream)
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
            b = self._private_copy(body_wxyz_xyz.contiguous().double(), stream, "_body_stage")
            code = self._lib.sas_set_link_poses_from_bodies(self._h, C.c_void_p(b.data_ptr()), b.numel() // (self.L * 7), L.SAS_MEM_DEVICE,
                                                            self._stream(stream))
            self._keep_bodies = b
            self._after_private_read(stream)
        else:
            b = np.ascontiguousarray(body_wxyz_xyz, dtype=np.float64).reshape(-1, self.L, 7)
            code = self._lib.sas_set_link_poses_from_bodies(self._h, b.ctypes.data, b.shape[0], L.SAS_MEM_HOST, self._stream(stream))
        L.check(self._lib, self._h, code)
        self._handback()

    def set_fk_model(self, source, n_bodies: Optional[int] = None) -> None:
        """URDF joint tree for the on-device forward kinematics (csrc/sas_pose.cu fk_kernel).  `source` is a
        urdf_fk.BatchedBodyPoseSource; its first n_bodies mesh-carrying bodies drive link groups 0..n_bodies-1 (default:
        as many as the URDF has, at most L), the remaining groups are free bodies (the T-block)."""
        nb = min(self.L, len(source.bodies)) if n_bodies is None else int(n_bodies)
        d = source.device_model(nb)
        self._fk_keep = d   # the arrays must outlive the call
        m = L.FkModel()
        m.n_joints, m.n_actuated, m.n_bodies = len(d["parent"]), int(d["n_actuated"]), nb
        for k in ("parent", "type", "qindex", "mimic", "origin", "axis", "base", "body_frame"):
            setattr(m, k, d[k].ctypes.data)
        L.check(self._lib, self._h, self._lib.sas_set_fk_model(self._h, C.byref(m)))
        self._fk_actuated, self._fk_bodies = m.n_actuated, nb

    def set_link_poses_from_joints(self, q, free_bodies=None, stream=None) -> None:
        """Joint angles [E, A] fp64 -> body poses (URDF forward kinematics) -> link poses, all on the device; needs
        set_fk_model and set_pose_model.  free_bodies [E, L - n_bodies, 7] fp64 (wxyz | xyz): poses of the link groups
        that are not robot bodies (default identity).  numpy (host) or CUDA float64 tensors."""
        torch = self._torch
        dev_in = isinstance(q, torch.Tensor) and q.is_cuda
        if dev_in:
            qq = self._private_copy(q.contiguous().double(), stream, "_q_stage")
            fb = None
            if free_bodies is not None:
                fb = self._private_copy(free_bodies.contiguous().double(), stream, "_fb_stage")
            E = qq.numel() // self._fk_actuated
            code = self._lib.sas_set_link_poses_from_joints(self._h, C.c_void_p(qq.data_ptr()), E, C.c_void_p(fb.data_ptr()) if fb is not None else None,
                                                            L.SAS_MEM_DEVICE, self._stream(stream))
            self._keep_q = (qq, fb)
            self._after_private_read(stream)
        else:
            qq = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self._fk_actuated)
            fb = None if free_bodies is None else np.ascontiguousarray(free_bodies, dtype=np.float64).reshape(qq.shape[0], -1, 7)
            code = self._lib.sas_set_link_poses_from_joints(self._h, qq.ctypes.data, qq.shape[0], fb.ctypes.data if fb is not None else None,
                                                            L.SAS_MEM_HOST, self._stream(stream))
            self._keep_q = (qq, fb)
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
        """viewmats [I,4,4] world->camera (OpenCV axes), Ks [I,3,3] or [3,3]; host arrays.  I may exceed the handle's
        max_images: render() then goes through the images in chunks of at most max_images (one library call each)."""
        vm = _f32c(viewmats).reshape(-1, 16)
        I = vm.shape[0]
        K = _f32c(Ks).reshape(-1, 9)
        if K.shape[0] == 1 and I > 1:
            K = np.ascontiguousarray(np.repeat(K, I, axis=0))
        env = None
        if env_of_image is not None:
            env = np.ascontiguousarray(env_of_image, dtype=np.int32)
            assert env.shape == (I,)
        self._cams = (vm, K, env)
        if (I, int(width), int(height)) != (self.I, self.W, self.H):
            self._sized = False   # another batch shape: the next unchecked render re-sizes the intersection buffers first
        self.I, self.W, self.H = I, int(width), int(height)
        self._chunk = None if I <= self.max_images else self.max_images   # images per library call in chunked mode
        if self._chunk is None:
            self._upload_cameras(0, I, stream)

    def _upload_cameras(self, i0: int, i1: int, stream=None) -> None:
        vm, K, env = self._cams
        e = env[i0:i1] if env is not None else None
        code = self._lib.sas_set_cameras(self._h, vm[i0:i1].ctypes.data, K[i0:i1].ctypes.data, e.ctypes.data if e is not None else None,
                                         i1 - i0, self.W, self.H, self._stream(stream))
        L.check(self._lib, self._h, code)
        self._handback()

    # -- render -----------------------------------------------------------------------------
    _SHAPES = {"rgb": (3, "float32"), "depth": (0, "float32"), "alpha": (0, "float32"), "rgb_u8": (3, "uint8"),
               "n_contrib": (0, "int32"), "depth_f16": (0, "float16")}

    def _shape_dtype(self, name: str):
        ch, dt = self._SHAPES[name]
        shape = (self.I, self.H, self.W, 3) if ch else (self.I, self.H, self.W)
        return shape, getattr(self._torch, dt)

    def _alloc(self, name: str, to_host: bool, slot: int = 0):
        torch = self._torch
        shape, dt = self._shape_dtype(name)
        if to_host:
            t = self._pinned.get((name, slot))
            if t is None or tuple(t.shape) != shape:
                t = torch.empty(shape, dtype=dt, pin_memory=True)
                self._pinned[(name, slot)] = t
            return t
        return torch.empty(shape, dtype=dt, device=f"cuda:{self.device}")

    def _alloc_packed(self, names: Sequence[str], slot: int = 0) -> dict:
        """Pinned host tensors for `names`, all views of ONE pinned block cut up by sas_packed_layout: the library
        then brings the whole observation back in a single D2H copy."""
        torch = self._torch
        key = ("packed", tuple(names), slot, self.I, self.H, self.W)
        hit = self._pinned.get(key)
        if hit is not None:
            return hit
        want = 0
        for n in names:
            want |= 1 << L.OUT_ORDER.index(n)
        offs = (C.c_size_t * 6)()
        total = C.c_size_t()
        code = self._lib.sas_packed_layout(self.I, self.W, self.H, want, offs, C.byref(total))
        if code != L.SAS_OK:
            raise L.SasError(code, "sas_packed_layout: bad shape or output mask")
        block = torch.empty(int(total.value), dtype=torch.uint8, pin_memory=True)
        bufs = {"_block": block}
        for n in names:
            shape, dt = self._shape_dtype(n)
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
            o = int(offs[L.OUT_ORDER.index(n)])
            bufs[n] = block[o:o + nbytes].view(dt).view(shape)
        for k in [k for k in self._pinned if isinstance(k, tuple) and k[0] == "packed" and k[2] == slot and k[1] == tuple(names) and k != key]:
            del self._pinned[k]   # a block of another shape for the same outputs
        self._pinned[key] = bufs
        return bufs

    def _outputs_struct(self, bufs: dict, mem: int, nerfstudio_post: bool, background, offset_images: int = 0) -> "L.Outputs":
        o = L.Outputs()
        o.mem = mem
        o.nerfstudio_post = 1 if nerfstudio_post else 0
        o.background = (C.c_float * 3)(*[float(b) for b in background])
        for n in L.OUT_ORDER:
            t = bufs.get(n)
            if t is None:
                setattr(o, n, None)
            else:
                per_image = t[0].numel() * t.element_size() if t.shape[0] else 0
                setattr(o, n, C.c_void_p(t.data_ptr() + offset_images * per_image))
        return o

    _ISECT_CAP = 0x7ffffff0   # int32 offsets inside the library

    def _render_checked(self, o, st) -> bool:
        """One library call; on overflow grow the intersection buffers (from the device counters) and render again.
        Returns False if the frame cannot fit the library's int32 offsets (the caller then halves the chunk)."""
        for attempt in range(3):
            try:
                L.check(self._lib, self._h, self._lib.sas_render(self._h, C.byref(o), st))
            except L.SasError as e:
                if e.code != -4:
                    raise
            s = self.stats()
            if not s["overflow"]:
                return True
            if s["n_isects"] > self._ISECT_CAP or s["n_isects_src"] > self._ISECT_CAP:
                return False
            need_src = int(s["n_isects_src"] * 1.25) + 1024 if s["n_isects_src"] > s["cap_isects_src"] else 0
            need_all = int(s["n_isects"] * 1.25) + 1024 if s["n_isects"] > s["cap_isects"] else 0
            L.check(self._lib, self._h, self._lib.sas_reserve_isects2(self._h, min(need_src, self._ISECT_CAP), min(need_all, self._ISECT_CAP)))
        raise L.SasError(-4, "intersection buffers still overflow after growing them twice")

    def render(self, outputs: Iterable[str] = ("rgb", "depth", "alpha"), background=(0.0, 0.0, 0.0),
               nerfstudio_post: bool = False, to_host: bool = False, out: Optional[dict] = None, stream=None,
               check: bool = True) -> dict:
        """Render every image set by set_cameras.  Returns {name: torch tensor}: CUDA tensors, or (to_host=True)
        pinned host tensors filled by a D2H copy inside the call (which then synchronises the stream).
        check=True reads the device counters afterwards (a sync) and transparently grows the intersection
        buffers and re-renders if they overflowed; check=False never synchronises -- overflow then surfaces as an
        exception from a host-output render, or through poll_overflow()."""
        if self.I == 0:
            raise RuntimeError("set_cameras has not been called")
        names = list(outputs)
        if self._chunk is not None:
            return self._render_chunked(names, background, nerfstudio_post, to_host, out, check)
        bufs = dict(out) if out else {}
        missing = [n for n in names if n not in bufs]
        if missing:
            if to_host and not out:
                bufs.update(self._alloc_packed(names))
            else:
                for n in missing:
                    bufs[n] = self._alloc(n, to_host)
        o = self._outputs_struct({n: bufs[n] for n in names}, L.SAS_MEM_HOST if to_host else L.SAS_MEM_DEVICE, nerfstudio_post, background)
        st = self._stream(stream)
        if check:
            if not self._render_checked(o, st):
                raise L.SasError(-4, "this camera set needs more than 2^31 intersections in one call: create the renderer with a "
                                     "smaller max_images so that render() walks the images in chunks")
            self._sized = True
        else:
            L.check(self._lib, self._h, self._lib.sas_render(self._h, C.byref(o), st))
        self._handback()
        return {n: bufs[n] for n in names}

    def _render_chunked(self, names, background, nerfstudio_post, to_host, out, check: bool = True) -> dict:
        """More images than one library call takes (or than fit its int32 intersection offsets): the images are
        rendered max_images at a time into slices of the full output tensors.  The image index is the most
        significant field of gsplat's sort key, so every chunk is exactly the corresponding part of the one-call
        result.  A chunk that still overflows the offsets is halved.  check=False (after a checked pass has sized the
        buffers and settled the chunk size) issues the calls without reading the counters back."""
        torch = self._torch
        dev = dict(out) if (out and not to_host) else {}
        for n in names:
            if n not in dev:
                dev[n] = self._alloc(n, False)
        i0, step = 0, int(self._chunk)
        while i0 < self.I:
            i1 = min(self.I, i0 + step)
            self._upload_cameras(i0, i1)
            o = self._outputs_struct({n: dev[n] for n in names}, L.SAS_MEM_DEVICE, nerfstudio_post, background, offset_images=i0)
            if not check:
                L.check(self._lib, self._h, self._lib.sas_render(self._h, C.byref(o), self._stream(None)))
            elif not self._render_checked(o, self._stream(None)):
                if step == 1:
                    raise L.SasError(-4, "a single image needs more than 2^31 intersections")
                step = max(1, step // 2)
                self._chunk = step   # remembered: later (unchecked) passes use the size that fits
                continue
            self._handback()
            i0 = i1
        if not to_host:
            return {n: dev[n] for n in names}
        host = dict(out) if out else {}
        for n in names:
            if n not in host:
                host[n] = self._alloc(n, True)
            host[n].copy_(dev[n], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return {n: host[n] for n in names}

    def _size_buffers_once(self, names, background, nerfstudio_post, stream) -> None:
        """Unchecked (asynchronous) renders cannot grow the intersection buffers themselves, so the first one after
        set_cameras is preceded by ONE checked device render with the same poses, which sizes the buffers (plus 50 %
        head-room for the steps to come).  Later overflows are reported, never silent: wait() raises for host
        outputs, poll_overflow() tells for device outputs."""
        if self._sized:
            return
        if self._chunk is not None:
            raise RuntimeError("asynchronous rendering is not available in chunked mode (more images than max_images)")
        dev = {n: self._alloc(n, False) for n in names}
        o = self._outputs_struct(dev, L.SAS_MEM_DEVICE, nerfstudio_post, background)
        if not self._render_checked(o, self._stream(stream)):
            raise L.SasError(-4, "this camera set needs more than 2^31 intersections in one call")
        s = self.stats()
        L.check(self._lib, self._h, self._lib.sas_reserve_isects2(self._h, min(int(s["n_isects_src"] * 1.5) + 4096, self._ISECT_CAP),
                                                                  min(int(s["n_isects"] * 1.5) + 4096, self._ISECT_CAP)))
        self._handback()
        self._sized = True

    def render_async(self, outputs: Iterable[str] = ("rgb_u8", "depth"), background=(0.0, 0.0, 0.0),
                     nerfstudio_post: bool = False, stream=None) -> int:
        """Issue a render into pinned host buffers WITHOUT waiting for it (sas_outputs.mem = SAS_MEM_HOST_ASYNC): the
        D2H copy (ONE, the outputs share a pinned block) runs on the library's copy stream out of one of two
        alternating staging sets, so the transfer of step k overlaps the render of step k+1.  Returns a ticket for
        wait().  At most two renders may be outstanding (issue k+1, then wait(k)); the tensors wait(k) returned are
        reused by render k+2.  The first call sizes the intersection buffers with one checked render; if a later step
        overflows them, wait() raises (the frame is incomplete) -- reserve_isects() and carry on."""
        if self.I == 0:
            raise RuntimeError("set_cameras has not been called")
        seq = getattr(self, "_async_seq", 0)
        names = list(outputs)
        self._size_buffers_once(names, background, nerfstudio_post, stream)
        bufs = self._alloc_packed(names, slot=seq & 1)
        o = self._outputs_struct({n: bufs[n] for n in names}, L.SAS_MEM_HOST_ASYNC, nerfstudio_post, background)
        L.check(self._lib, self._h, self._lib.sas_render(self._h, C.byref(o), self._stream(stream)))
        self._handback()
        if not hasattr(self, "_async_bufs"):
            self._async_bufs = [None, None]
        self._async_bufs[seq & 1] = {n: bufs[n] for n in names}
        self._async_seq = seq + 1
        return seq

    def wait(self, ticket: Optional[int] = None) -> dict:
        """Block until the host buffers of render_async ticket (default: the latest) are complete; returns them.
        Raises SasError(-4) if that frame overflowed the intersection buffers (it is incomplete)."""
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

    def reserve_isects(self, n: int, n_src: Optional[int] = None) -> None:
        """Capacity for n merged intersections and n_src sorted bucket entries (default: n)."""
        L.check(self._lib, self._h, self._lib.sas_reserve_isects2(self._h, int(n if n_src is None else n_src), int(n)))

    def poll_overflow(self, stream=None) -> Optional[dict]:
        """Did one of the last (up to 4) renders since the previous poll overflow the intersection buffers?  Synchronises
        the stream.  None, or {'n_isects_src': ..., 'n_isects': ...} = the capacities that would have sufficed."""
        ns, na = C.c_longlong(), C.c_longlong()
        code = self._lib.sas_poll_overflow(self._h, C.byref(ns), C.byref(na), self._stream(stream))
        self._handback()
        if code < 0:
            L.check(self._lib, self._h, code)
        return {"n_isects_src": int(ns.value), "n_isects": int(na.value)} if code == 1 else None

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

    def set_fk_model(self, source, n_bodies: Optional[int] = None) -> None:
        for r in self.renderers:
            r.set_fk_model(source, n_bodies)

    def set_link_poses_from_joints(self, q, free_bodies=None) -> None:
        r, s = self._cur()
        self._after_caller(s)
        r.set_link_poses_from_joints(q, free_bodies, stream=s)

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
        r._size_buffers_once(list(names), background, nerfstudio_post, s)   # first use: one checked render sizes the buffers
        r.render(names, background=background, nerfstudio_post=nerfstudio_post, out=bufs, stream=s, check=False)
        self._seq += 1
        return bufs

    def poll_overflow(self) -> None:
        """Synchronises every handle's stream and raises SasError(-4) if one of their recent device-output renders
        overflowed its intersection buffers (those frames were incomplete); grows the buffers for the steps to come."""
        bad = None
        for r, s in zip(self.renderers, self.streams):
            need = r.poll_overflow(stream=s)
            if need:
                r.reserve_isects(int(need["n_isects"] * 1.5) + 4096, int(need["n_isects_src"] * 1.5) + 4096)
                bad = need
        if bad:
            raise L.SasError(-4, f"a recent render overflowed the intersection buffers (needed {bad}); they have been grown, "
                                 "the affected frames were incomplete")

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
