"""The C-ABI library builds here (nvcc cross-compiles sm_100a without a GPU), loads, exports every
symbol include/sas.h declares, and fails loudly -- not silently on a CPU path -- without a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from sim_a_splat_b200 import _build, _lib

ROOT = Path(__file__).resolve().parents[1]


def declared_functions():
    text = (ROOT / "include" / "sas.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sas_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    _build.build_library()
    lib = C.CDLL(str(_build.LIB))
    names = declared_functions()
    assert len(names) >= 11
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sas.h but not exported"
    assert sorted(_lib.EXPORTS) == names
    lib.sas_version.restype = C.c_char_p
    assert b"sm_100a" in lib.sas_version()


def test_sass_is_sm100a_and_uses_async_copies():
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", str(_build.LIB)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", str(_build.LIB)], capture_output=True, text=True).stdout
    assert "LDGSTS" in sass  # cp.async staging in the compositor
    assert "VOTE" in sass and "ATOMS" in sass  # ballot multi-split / shared-memory histograms in the bucket sort


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    from sim_a_splat_b200 import scene as S
    from sim_a_splat_b200.renderer import SplatRenderer

    sc = S.make_synthetic_scene(64, 0, seed=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SplatRenderer(sc)
    from sim_a_splat_b200.renderer import PipelinedSplatRenderer

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PipelinedSplatRenderer(sc, depth=2)  # the two-steps-in-flight front end has no CPU path either
    # and straight through the C ABI
    lib = _lib.load()
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (sc.means, sc.quats, sc.scales, sc.opacities, sc.sh)]
    lid = np.ascontiguousarray(sc.link_id, dtype=np.int32)
    desc = _lib.SceneDesc(sc.N, 0, 0, _lib.SAS_MEM_HOST, *[a.ctypes.data for a in arrs], lid.ctypes.data)
    cfg = _lib.Config(1, 1, 64, 64, 0)
    h = C.c_void_p()
    rc = lib.sas_create(C.byref(desc), C.byref(cfg), 0, C.byref(h))
    assert rc == -2 and not h.value  # SAS_ECUDA
    assert b"no CPU fallback" in lib.sas_last_error(None)


def test_entry_points_reject_null_handles_without_touching_a_device():
    lib = _lib.load()
    assert lib.sas_wait_outputs(None, 0) == -1          # SAS_EINVAL
    assert lib.sas_render(None, None, None) == -1
    assert lib.sas_last_launch_count(None) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_numa_binding_is_a_no_op_without_gpu():
    import os

    from sim_a_splat_b200.dist import bind_to_gpu_numa

    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa(0) is None
    assert os.sched_getaffinity(0) == before


def test_create_rejects_bad_arguments():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.sas_create(None, None, 0, C.byref(h)) == -1
    desc = _lib.SceneDesc(0, 0, 0, 0, None, None, None, None, None, None)
    cfg = _lib.Config(1, 1, 64, 64, 0)
    assert lib.sas_create(C.byref(desc), C.byref(cfg), 0, C.byref(h)) == -1
    assert b"bad scene" in lib.sas_last_error(None)
