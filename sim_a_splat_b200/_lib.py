"""ctypes binding of libsas_b200.so -- the C ABI declared in include/sas.h.

The product path has no CPU fallback: if the shared library is missing it is
built in-tree with nvcc (sim_a_splat_b200/_build.py); if that fails, or no CUDA
device is usable, the failure is raised, never papered over.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from . import _build

SAS_OK = 0
SAS_MEM_HOST = 0
SAS_MEM_DEVICE = 1
SAS_MEM_HOST_ASYNC = 2
SAS_DBG_RECORDS = 1
SAS_DBG_BINS = 2
SAS_DBG_TILE_OFFSETS = 3
SAS_DBG_SORTED_GIDS = 4
SAS_DBG_VIEW_OF_IMAGE = 5
SAS_DBG_RADII = 6
SAS_DBG_LINK_POSES = 7

EXPORTS = [
    "sas_create", "sas_destroy", "sas_set_link_poses", "sas_set_cameras", "sas_render", "sas_reserve_isects",
    "sas_set_pose_model", "sas_set_link_poses_from_bodies", "sas_unproject_depth", "sas_wait_outputs",
    "sas_get_stats", "sas_set_stage_timing", "sas_get_stage_ms", "sas_debug_fetch", "sas_last_launch_count", "sas_last_error", "sas_version",
    "sas_reserve_isects2", "sas_packed_layout", "sas_poll_overflow", "sas_set_fk_model", "sas_set_link_poses_from_joints",
]

# bit k of a SAS_OUT_* mask <-> output name (the order of sas_packed_layout)
OUT_ORDER = ("rgb_u8", "depth_f16", "depth", "alpha", "rgb", "n_contrib")


class SceneDesc(C.Structure):
    _fields_ = [("n_gaussians", C.c_int32), ("sh_degree", C.c_int32), ("n_links", C.c_int32), ("mem", C.c_int32),
                ("means", C.c_void_p), ("quats", C.c_void_p), ("scales", C.c_void_p), ("opacities", C.c_void_p),
                ("sh", C.c_void_p), ("link_id", C.c_void_p)]


class Config(C.Structure):
    _fields_ = [("max_images", C.c_int32), ("max_envs", C.c_int32), ("max_width", C.c_int32), ("max_height", C.c_int32),
                ("max_isects", C.c_int64), ("max_isects_src", C.c_int64), ("radius_rule", C.c_int32), ("reserved_", C.c_int32)]


class Outputs(C.Structure):
    _fields_ = [("mem", C.c_int32), ("nerfstudio_post", C.c_int32), ("background", C.c_float * 3),
                ("rgb", C.c_void_p), ("depth", C.c_void_p), ("alpha", C.c_void_p), ("rgb_u8", C.c_void_p),
                ("n_contrib", C.c_void_p), ("depth_f16", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("n_isects", C.c_int64), ("n_isects_src", C.c_int64), ("n_visible_src", C.c_int64), ("max_tile_list", C.c_int64),
                ("n_staged", C.c_int64), ("overflow", C.c_int32), ("n_images", C.c_int32), ("n_views", C.c_int32),
                ("n_gaussians", C.c_int32), ("n_background", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("tile_w", C.c_int32), ("tile_h", C.c_int32), ("pad_", C.c_int32), ("cap_isects", C.c_int64),
                ("cap_isects_src", C.c_int64)]


class PoseModel(C.Structure):
    _fields_ = [("n_links", C.c_int32), ("scale", C.c_double), ("Ri", C.c_double * 9), ("ti", C.c_double * 3),
                ("fk_R", C.c_void_p), ("fk_t", C.c_void_p), ("body_offset", C.c_void_p)]


class FkModel(C.Structure):
    _fields_ = [("n_joints", C.c_int32), ("n_actuated", C.c_int32), ("n_bodies", C.c_int32), ("parent", C.c_void_p), ("type", C.c_void_p),
                ("qindex", C.c_void_p), ("mimic", C.c_void_p), ("origin", C.c_void_p), ("axis", C.c_void_p), ("base", C.c_void_p),
                ("body_frame", C.c_void_p)]


class SasError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsas_b200 error {code}: {msg}")
        self.code = code


_lib = None


def library_path() -> Path:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Load (building first if needed) libsas_b200.so and declare the C ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing and _build.needs_build():
        _build.build_library()
    if not _build.LIB.exists():
        raise RuntimeError(f"{_build.LIB} is missing and could not be built; sim_a_splat_b200 has no CPU fallback")
    lib = C.CDLL(str(_build.LIB))
    P = C.c_void_p
    lib.sas_create.argtypes = [C.POINTER(SceneDesc), C.POINTER(Config), C.c_int, C.POINTER(P)]
    lib.sas_create.restype = C.c_int
    lib.sas_destroy.argtypes = [P]
    lib.sas_destroy.restype = None
    lib.sas_set_link_poses.argtypes = [P, P, C.c_int, C.c_int, P]
    lib.sas_set_link_poses.restype = C.c_int
    lib.sas_set_pose_model.argtypes = [P, C.POINTER(PoseModel)]
    lib.sas_set_pose_model.restype = C.c_int
    lib.sas_set_link_poses_from_bodies.argtypes = [P, P, C.c_int, C.c_int, P]
    lib.sas_set_link_poses_from_bodies.restype = C.c_int
    lib.sas_set_fk_model.argtypes = [P, C.POINTER(FkModel)]
    lib.sas_set_fk_model.restype = C.c_int
    lib.sas_set_link_poses_from_joints.argtypes = [P, P, C.c_int, P, C.c_int, P]
    lib.sas_set_link_poses_from_joints.restype = C.c_int
    lib.sas_unproject_depth.argtypes = [P, P, C.c_int, C.c_int, C.c_int, P, C.c_float, P, P, C.c_int, P]
    lib.sas_unproject_depth.restype = C.c_int
    lib.sas_set_cameras.argtypes = [P, P, P, P, C.c_int, C.c_int, C.c_int, P]
    lib.sas_set_cameras.restype = C.c_int
    lib.sas_render.argtypes = [P, C.POINTER(Outputs), P]
    lib.sas_render.restype = C.c_int
    lib.sas_wait_outputs.argtypes = [P, C.c_int]
    lib.sas_wait_outputs.restype = C.c_int
    lib.sas_reserve_isects.argtypes = [P, C.c_longlong]
    lib.sas_reserve_isects.restype = C.c_int
    lib.sas_reserve_isects2.argtypes = [P, C.c_longlong, C.c_longlong]
    lib.sas_reserve_isects2.restype = C.c_int
    lib.sas_packed_layout.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.sas_packed_layout.restype = C.c_int
    lib.sas_poll_overflow.argtypes = [P, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), P]
    lib.sas_poll_overflow.restype = C.c_int
    lib.sas_get_stats.argtypes = [P, C.POINTER(Stats)]
    lib.sas_get_stats.restype = C.c_int
    lib.sas_set_stage_timing.argtypes = [P, C.c_int]
    lib.sas_set_stage_timing.restype = C.c_int
    lib.sas_get_stage_ms.argtypes = [P, C.POINTER(C.c_float)]
    lib.sas_get_stage_ms.restype = C.c_int
    lib.sas_debug_fetch.argtypes = [P, C.c_int, P, C.c_size_t]
    lib.sas_debug_fetch.restype = C.c_int
    lib.sas_last_launch_count.argtypes = [P]
    lib.sas_last_launch_count.restype = C.c_int
    lib.sas_last_error.argtypes = [P]
    lib.sas_last_error.restype = C.c_char_p
    lib.sas_version.argtypes = []
    lib.sas_version.restype = C.c_char_p
    _lib = lib
    return lib


def check(lib, handle, code: int) -> None:
    if code != SAS_OK:
        msg = lib.sas_last_error(handle)
        raise SasError(code, msg.decode() if msg else "")
