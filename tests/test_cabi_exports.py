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
    # the bulk-async staging variant of the compositor (SAS_VARIANT bit 11: cp.async.bulk + mbarrier expect_tx) is in the
    # binary too -- measured slower than LDGSTS for this 48-byte gather, profiles/r02b_bulk_staging.txt
    assert "UBLKCP" in sass and "SYNCS" in sass


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
    assert lib.sas_reserve_isects2(None, 1, 1) == -1
    assert lib.sas_poll_overflow(None, None, None, None) == -1
    assert lib.sas_set_fk_model(None, None) == -1
    assert lib.sas_set_link_poses_from_joints(None, None, 1, None, 0, None) == -1


def test_packed_layout_is_pure_host_arithmetic():
    """sas_packed_layout: the one-block layout of the host outputs (no device involved): outputs in SAS_OUT_* bit order, every
    offset 256-byte aligned, unwanted outputs untouched."""
    lib = _lib.load()
    offs = (C.c_size_t * 6)(*([12345] * 6))
    total = C.c_size_t()
    I, W, H = 256, 96, 96
    P = I * W * H
    want = 1 | 4                                          # uint8 RGB + f32 depth: the bench's end-to-end observation
    assert lib.sas_packed_layout(I, W, H, want, offs, C.byref(total)) == 0
    assert offs[0] == 0 and offs[2] == ((3 * P + 255) // 256) * 256 and offs[1] == 12345 and offs[3] == 12345
    assert total.value == offs[2] + ((4 * P + 255) // 256) * 256
    assert lib.sas_packed_layout(I, W, H, 63, offs, C.byref(total)) == 0
    sizes = [3, 2, 4, 4, 12, 4]
    for k in range(6):
        assert offs[k] % 256 == 0
        if k:
            assert offs[k] >= offs[k - 1] + sizes[k - 1] * P
    assert total.value >= offs[5] + 4 * P
    assert lib.sas_packed_layout(0, W, H, 1, offs, C.byref(total)) == -1
    assert lib.sas_packed_layout(I, W, H, 64, offs, C.byref(total)) == -1


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
