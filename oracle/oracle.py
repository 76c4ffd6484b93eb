"""ctypes binding of oracle/splat_oracle.c (CPU ORACLE -- test infrastructure, see its header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build_oracle(force: bool = False) -> Path:
    """Compile liboracle.so with the recipe in oracle/Makefile (gcc, OpenMP, -ffp-contract=off)."""
    so = _HERE / "liboracle.so"
    src = _HERE / "splat_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "liboracle.so"], check=True, capture_output=True)
    return so


class _Scene(C.Structure):
    _fields_ = [("N", C.c_int32), ("sh_degree", C.c_int32), ("L", C.c_int32),
                ("means", C.c_void_p), ("quats", C.c_void_p), ("scales", C.c_void_p),
                ("opacities", C.c_void_p), ("sh", C.c_void_p), ("link_id", C.c_void_p)]


class _Frame(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("I", "N", "W", "H", "tile_w", "tile_h", "tile_bits", "image_bits")] + \
               [("M", C.c_int64)] + \
               [(n, C.c_void_p) for n in ("radii", "means2d", "depths", "conics", "colors", "tiles_per_gauss",
                                          "tile_rect", "isect_ids", "flatten_ids", "isect_offsets", "rgb",
                                          "depth", "alpha", "n_contrib", "last_idx")] + \
               [("pairs_evaluated", C.c_int64)]


def _lib():
    global _LIB
    if _LIB is None:
        lib = C.CDLL(str(build_oracle()))
        lib.orc_render.restype = C.POINTER(_Frame)
        lib.orc_render.argtypes = [C.POINTER(_Scene), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        lib.orc_render2.restype = C.POINTER(_Frame)
        lib.orc_render2.argtypes = lib.orc_render.argtypes + [C.c_int]
        lib.orc_extend_export.restype = C.c_float
        lib.orc_extend_export.argtypes = [C.c_int, C.c_float]
        lib.orc_frame_free.argtypes = [C.POINTER(_Frame)]
        lib.orc_frame_free.restype = None
        lib.orc_exp_neg_export.restype = C.c_float
        lib.orc_exp_neg_export.argtypes = [C.c_float]
        lib.orc_compute_cov.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.orc_compute_cov.restype = None
        lib.orc_num_threads.restype = C.c_int
        lib.orc_set_num_threads.argtypes = [C.c_int]
        _LIB = lib
    return _LIB


def num_threads() -> int:
    return int(_lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    _lib().orc_set_num_threads(int(n))


def exp_neg(x: float) -> float:
    return float(_lib().orc_exp_neg_export(float(x)))


def compute_cov(quats: np.ndarray, scales: np.ndarray) -> np.ndarray:
    q = np.ascontiguousarray(quats, dtype=np.float32)
    s = np.ascontiguousarray(scales, dtype=np.float32)
    out = np.empty((q.shape[0], 3, 3), dtype=np.float32)
    _lib().orc_compute_cov(q.ctypes.data, s.ctypes.data, q.shape[0], out.ctypes.data)
    return out


class OracleFrame(dict):
    """dict of numpy arrays (copies) + scalar metadata."""

    __getattr__ = dict.__getitem__


def camera_origin(viewmats: np.ndarray) -> np.ndarray:
    """Camera origin in world, -Rc^T tc, computed in fp64 and rounded once to fp32.
    Both the oracle and the product receive this same fp32 array."""
    vm = np.asarray(viewmats, dtype=np.float64).reshape(-1, 4, 4)
    return (-np.einsum("nji,nj->ni", vm[:, :3, :3], vm[:, :3, 3])).astype(np.float32)


RADIUS_RULES = {"torch": 0, "cuda_1_5": 1}


def render(scene: dict, viewmats, Ks, W: int, H: int, link_poses=None, env_of_image=None,
           bg=(0.0, 0.0, 0.0), stages: int = 3, keep=None, radius_rule: str = "torch") -> OracleFrame:
    """scene: dict(means[N,3], quats[N,4] wxyz, scales[N,3] linear, opacities[N] linear,
    sh[N,K,3], link_id[N] int32 (-1 background), L).  link_poses [E,L,12] fp32."""
    lib = _lib()
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)  # noqa: E731
    means, quats, scales = f32(scene["means"]), f32(scene["quats"]), f32(scene["scales"])
    opac, sh = f32(scene["opacities"]).reshape(-1), f32(scene["sh"])
    N = means.shape[0]
    K = sh.shape[1]
    deg = int(round(np.sqrt(K))) - 1
    assert (deg + 1) ** 2 == K and sh.shape == (N, K, 3)
    link_id = np.ascontiguousarray(scene.get("link_id", -np.ones(N)), dtype=np.int32)
    L = int(scene.get("L", 0))
    vm = f32(viewmats).reshape(-1, 16)
    I = vm.shape[0]
    Kc = f32(Ks).reshape(I, 9)
    campos = camera_origin(vm.reshape(I, 4, 4))
    if env_of_image is None:
        env_of_image = np.zeros(I)
    env = np.ascontiguousarray(env_of_image, dtype=np.int32)
    if L > 0:
        lp = f32(link_poses).reshape(-1, L, 12)
        E = lp.shape[0]
        assert env.max() < E
    else:
        lp, E = None, 0
    bgv = f32(bg)
    sc = _Scene(N, deg, L, means.ctypes.data, quats.ctypes.data, scales.ctypes.data, opac.ctypes.data,
                sh.ctypes.data, link_id.ctypes.data)
    fp = lib.orc_render2(C.byref(sc), lp.ctypes.data if lp is not None else None, E, vm.ctypes.data,
                         Kc.ctypes.data, campos.ctypes.data, env.ctypes.data, I, W, H, bgv.ctypes.data, stages, RADIUS_RULES[radius_rule])
    f = fp.contents
    M = int(f.M)
    th, tw = f.tile_h, f.tile_w

    def arr(ptr, shape, dt, name=None):
        if not ptr or (keep is not None and name not in keep):
            return None
        n = int(np.prod(shape))
        if n == 0:
            return np.zeros(shape, dtype=dt)
        buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dt).reshape(shape).copy()

    out = OracleFrame(
        I=I, N=N, W=W, H=H, tile_w=tw, tile_h=th, tile_bits=f.tile_bits, image_bits=f.image_bits, M=M,
        pairs_evaluated=int(f.pairs_evaluated), campos=campos,
        radii=arr(f.radii, (I, N, 2), np.int32, "radii"), means2d=arr(f.means2d, (I, N, 2), np.float32, "means2d"),
        depths=arr(f.depths, (I, N), np.float32, "depths"), conics=arr(f.conics, (I, N, 3), np.float32, "conics"),
        colors=arr(f.colors, (I, N, 3), np.float32, "colors"), tiles_per_gauss=arr(f.tiles_per_gauss, (I, N), np.int32, "tiles_per_gauss"),
        tile_rect=arr(f.tile_rect, (I, N, 4), np.int32, "tile_rect"),
        isect_ids=arr(f.isect_ids, (M,), np.int64, "isect_ids"), flatten_ids=arr(f.flatten_ids, (M,), np.int32, "flatten_ids"),
        isect_offsets=arr(f.isect_offsets, (I, th, tw), np.int32, "isect_offsets"),
        rgb=arr(f.rgb, (I, H, W, 3), np.float32, "rgb"), depth=arr(f.depth, (I, H, W), np.float32, "depth"),
        alpha=arr(f.alpha, (I, H, W), np.float32, "alpha"), n_contrib=arr(f.n_contrib, (I, H, W), np.int32, "n_contrib"),
        last_idx=arr(f.last_idx, (I, H, W), np.int32, "last_idx"),
    )
    lib.orc_frame_free(fp)
    return out


def nerfstudio_outputs(frame: OracleFrame, background=(0.1490, 0.1647, 0.2157)) -> dict:
    """nerfstudio splatfacto `get_outputs` post-processing [RECALLED] on a frame rendered with bg=0:
    rgb = clamp(render + (1-alpha)*background, 0, 1); depth = where(alpha>0, ED, max(ED))."""
    bgv = np.asarray(background, dtype=np.float32)
    a = frame["alpha"][..., None]
    rgb = np.clip(frame["rgb"] + (np.float32(1.0) - a) * bgv, 0.0, 1.0).astype(np.float32)
    d = frame["depth"].copy()
    for i in range(d.shape[0]):
        d[i] = np.where(frame["alpha"][i] > 0, d[i], d[i].max())
    return {"rgb": rgb, "depth": d[..., None], "accumulation": a}
