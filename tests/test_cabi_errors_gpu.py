"""Error behaviour at the C ABI (include/sas.h), on a GPU box: negative codes + text, never a silent success."""
import numpy as np
import pytest

from sim_a_splat_b200 import scene as S

pytestmark = pytest.mark.gpu


def test_cabi_error_codes_and_messages():
    """Error behaviour at the C ABI (include/sas.h): negative codes + text, never a silent success."""
    import ctypes as C

    from sim_a_splat_b200 import _lib as L

    lib = L.load()
    sc = S.make_synthetic_scene(500, 1, seed=3, link_counts=(40,))
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (sc.means, sc.quats, sc.scales, sc.opacities, sc.sh)]
    lid = np.ascontiguousarray(sc.link_id, dtype=np.int32)
    desc = L.SceneDesc(sc.N, 1, 1, L.SAS_MEM_HOST, *[a.ctypes.data for a in arrs], lid.ctypes.data)
    h = C.c_void_p()
    # a scene that is not in canonical order (a link Gaussian before a background one) is refused
    bad = lid.copy()
    bad[0], bad[-1] = 0, -1
    bad_desc = L.SceneDesc(sc.N, 1, 1, L.SAS_MEM_HOST, *[a.ctypes.data for a in arrs], bad.ctypes.data)
    assert lib.sas_create(C.byref(bad_desc), C.byref(L.Config(2, 2, 64, 64, 0)), 0, C.byref(h)) == -1 and not h.value
    assert b"canonical" in lib.sas_last_error(None)
    assert lib.sas_create(C.byref(desc), C.byref(L.Config(2, 2, 64, 64, 0)), 99, C.byref(h)) == -2       # no such device
    assert lib.sas_create(C.byref(desc), C.byref(L.Config(2, 2, 5000, 64, 0)), 0, C.byref(h)) == -1     # width beyond 4080
    assert lib.sas_create(C.byref(desc), C.byref(L.Config(2, 2, 64, 64, 0)), 0, C.byref(h)) == 0 and h.value
    out = L.Outputs()
    out.mem = L.SAS_MEM_DEVICE
    assert lib.sas_render(h, C.byref(out), None) == -5 and b"sas_set_cameras" in lib.sas_last_error(h)  # call order
    vm = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (3, 1))
    K = np.tile(np.float32([50, 0, 32, 0, 50, 32, 0, 0, 1]), (3, 1))
    assert lib.sas_set_cameras(h, vm.ctypes.data, K.ctypes.data, None, 3, 64, 64, None) == -4           # max_images = 2
    assert lib.sas_set_cameras(h, vm.ctypes.data, K.ctypes.data, None, 2, 128, 64, None) == -4          # max_width = 64
    env = np.int32([0, 7])
    assert lib.sas_set_cameras(h, vm.ctypes.data, K.ctypes.data, env.ctypes.data, 2, 64, 64, None) == -1  # env out of range
    assert lib.sas_set_cameras(h, vm.ctypes.data, K.ctypes.data, None, 2, 64, 64, None) == 0
    assert lib.sas_render(h, C.byref(out), None) == -5 and b"sas_set_link_poses" in lib.sas_last_error(h)
    poses = np.tile(np.float32([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0]), (3, 1, 1))
    assert lib.sas_set_link_poses(h, poses.ctypes.data, 3, L.SAS_MEM_HOST, None) == -4                   # max_envs = 2
    assert lib.sas_set_link_poses(h, poses.ctypes.data, 2, L.SAS_MEM_HOST, None) == 0
    body = np.zeros((2, 1, 7))
    assert lib.sas_set_link_poses_from_bodies(h, body.ctypes.data, 2, L.SAS_MEM_HOST, None) == -5        # no pose model yet
    assert lib.sas_render(h, C.byref(out), None) == 0                                                    # no outputs requested: still valid
    st = L.Stats()
    assert lib.sas_get_stats(h, C.byref(st)) == 0 and st.n_images == 2 and st.overflow == 0
    buf = np.empty(10, dtype=np.int32)
    assert lib.sas_debug_fetch(h, 99, buf.ctypes.data, 4) == -1
    assert lib.sas_debug_fetch(h, L.SAS_DBG_TILE_OFFSETS, buf.ctypes.data, 10**9) == -1                  # more than available
    lib.sas_destroy(h)
