"""The reference-facing objects end to end on the GPU: SplatEnvWrapper.step() observations, the batched
VecSplatRenderer, HeadlessViser.get_render and GaussianSplat.render, each compared with the CPU oracle
rendering the same scene / link poses / cameras.  Reads like the reference's own scripts
(/root/reference/tests/test_splat_wrapper.py:63-131) with assertions added."""
import numpy as np
import pytest
import torch

import oracle
from sim_a_splat_b200 import dist as D
from sim_a_splat_b200 import scene as S
from sim_a_splat_b200.env.splat.splat_env_wrapper import SplatEnvWrapper
from sim_a_splat_b200.env.splat.vec_splat_env import VecSplatRenderer
from sim_a_splat_b200.ns_utils.nerfstudio_utils import GaussianSplat
from sim_a_splat_b200.renderer import NS_BACKGROUND

from compat_fixtures import N_LINKS, FakeManipulatorEnv, camera_setup, make_assets

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def toy(tmp_path_factory, assets):
    return make_assets(tmp_path_factory.mktemp("toy_assets_gpu"), assets)


def _to_u8(rgb):
    return np.floor(np.clip(rgb, 0.0, 1.0) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)


def _K(H, W, fov):
    fy = H / (2.0 * np.tan(fov / 2.0))
    return np.array([[fy, 0, W / 2.0], [0, fy, H / 2.0], [0, 0, 1.0]], dtype=np.float32)


def _oracle_views(env: SplatEnvWrapper, H, W):
    """What the oracle says the wrapper's cameras see right now."""
    srv = env.splat_handler.server
    sc = srv.splat_scene
    poses = env.get_moving_camera_poses(env.draw_msg) + env.fixed_cam_poses
    vms = np.stack([S.c2w_opencv_to_viewmat(p.rotation().wxyz, p.translation()) for p in poses])
    Ks = np.repeat(_K(H, W, env.ch.camera.fov)[None], len(poses), 0)
    return oracle.render(sc.as_oracle_dict(), vms, Ks, W, H, srv.link_poses(), np.zeros(len(poses), np.int32), bg=srv.background_rgb)


def test_env_wrapper_step_observation_matches_oracle(toy):
    H, W = 64, 80
    env = SplatEnvWrapper(FakeManipulatorEnv(toy, seed=3), toy["root"], toy["obj"], toy["cfg"], task_assets_name="tblock_paper.obj")
    env._configure_cameras(camera_setup((H, W)))
    obs = env.reset(seed=5)
    assert "camera_0" not in obs  # no image on reset (splat_env_wrapper.py:103)
    first = None
    for k in range(3):
        obs, reward, term, trunc, info = env.step(np.array([0.15, -0.1, 0.01]))
        assert set(obs) == {"joint_pos", "joint_vel", "camera_0", "camera_1"}
        ref = _oracle_views(env, H, W)
        assert ref["M"] > 500
        for ci in range(2):
            img = obs[f"camera_{ci}"]
            assert img.shape == (3, H, W) and img.dtype == np.uint8          # CHW uint8 (:135-137)
            want = np.moveaxis(_to_u8(ref["rgb"][ci]), -1, 0)
            assert np.sum(img != want) == 0, (k, ci)
        if first is None:
            first = {k_: v.copy() for k_, v in obs.items()}
    assert not np.array_equal(first["camera_1"], obs["camera_1"])  # the moving camera follows the gripper
    assert not np.array_equal(first["camera_0"], obs["camera_0"])  # the robot moved in the fixed view
    # render() alone and the client's get_render agree with the observation
    imgs = env.render()
    assert np.array_equal(np.moveaxis(imgs[1], -1, 0), obs["camera_1"])
    p = env.fixed_cam_poses[0]
    one = env.ch.get_render(height=H, width=W, wxyz=p.rotation().wxyz, position=p.translation())
    assert one.shape == (H, W, 3) and np.array_equal(np.moveaxis(one, -1, 0), obs["camera_1"])
    env.close()


def test_vec_renderer_matches_single_env_wrappers(toy):
    """E parallel envs through one render call == E single-env wrappers, image for image; sharded over 2 ranks
    (sim_a_splat_b200/dist.py) it is still the same set of images."""
    H, W, E = 48, 64, 5
    cams = camera_setup((H, W))
    singles, msgs = [], []
    for e in range(E):
        env = SplatEnvWrapper(FakeManipulatorEnv(toy, seed=100 + e), toy["root"], toy["obj"], toy["cfg"], task_assets_name="tblock_paper.obj")
        env._configure_cameras(cams)
        env.reset(seed=100 + e)
        obs, *_ = env.step(np.array([0.1 * e, -0.05 * e, 0.004 * e]))
        singles.append(obs)
        msgs.append(env.draw_msg)
        if e < E - 1:
            env.close()
    handler = env.splat_handler
    vec = VecSplatRenderer(handler, cams, n_envs=E, fov=env.ch.camera.fov, background=handler.server.background_rgb)
    out = vec.render(msgs)
    assert vec.renderer.stats()["n_views"] == 1 + E  # the fixed camera is one shared view, the moving one is per env
    for ci in range(2):
        assert out[f"camera_{ci}"].shape == (E, 3, H, W)
        for e in range(E):
            assert np.array_equal(out[f"camera_{ci}"][e], singles[e][f"camera_{ci}"]), (ci, e)
    vec.close()
    # two steps in flight on two handles (render_async / render_wait): same images, whichever handle drew them
    pv = VecSplatRenderer(handler, cams, n_envs=E, fov=env.ch.camera.fov, background=handler.server.background_rgb, pipeline_depth=2)
    t0 = pv.render_async(msgs)
    t1 = pv.render_async(msgs[::-1])
    a, t2 = pv.render_wait(t0), pv.render_async(msgs)
    for ci in range(2):
        assert np.array_equal(a[f"camera_{ci}"], out[f"camera_{ci}"]), ci
        assert np.array_equal(pv.render_wait(t1)[f"camera_{ci}"], out[f"camera_{ci}"][::-1]), ci
        assert np.array_equal(pv.render_wait(t2)[f"camera_{ci}"], out[f"camera_{ci}"]), ci
    with pytest.raises(RuntimeError):
        pv.render(msgs)
    pv.close()
    parts = []
    for rank in range(2):
        rng = D.shard_envs(E, 2, rank)
        v = VecSplatRenderer(handler, cams, n_envs=E, fov=env.ch.camera.fov, background=handler.server.background_rgb, env_range=rng)
        parts.append(v.render([msgs[e] for e in rng])["camera_0"])
        v.close()
    assert np.array_equal(np.concatenate(parts), out["camera_0"])
    env.close()


def test_task_block_is_rendered_and_follows_its_pose(toy):
    """the PushT T-block (a viser mesh in the reference, splat_handler.py:199-219) is drawn as a link
    group of surface Gaussians: it shows up in the observation in its red mesh colour, follows the block pose, and
    the batched path (device-side pose composition with the weld-free task link) gives the same images."""
    H, W = 72, 96
    kw = dict(task_assets_path=toy["task_assets_path"], task_assets_name=toy["task_assets_name"])
    from sim_a_splat_b200 import transforms as tf

    env = SplatEnvWrapper(FakeManipulatorEnv(toy, seed=3), toy["root"], toy["obj"], toy["cfg"], **kw)
    h = env.splat_handler
    # a camera 0.3 splat units above the block's work area, looking at it (block positions below are metric, URDF frame)
    target = h.scale_factor * (h.Ri @ np.array([0.07, 0.05, 0.02])) + h.ti
    eye = target + 0.3 * (h.Ri @ (np.array([0.25, -0.35, 0.9]) / np.linalg.norm([0.25, -0.35, 0.9])))
    cams = {0: dict(camera_setup((H, W))[0])}
    cams[0]["local_frame"] = tf.SE3.from_matrix(np.linalg.inv(S.lookat_viewmat(eye, target, up=h.Ri[:, 2]).astype(np.float64)))
    env._configure_cameras(cams)
    env.reset(seed=5)
    assert env.splat_handler.n_link_groups == N_LINKS + 1
    frames, msgs = [], []
    for block in ([0.0, 0.0, 0.0, 0.0], [0.15, 0.1, 0.0, 1.0]):
        env.unwrapped.block[:] = block
        obs, *_ = env.step(np.zeros(3))
        ref = _oracle_views(env, H, W)
        want = np.moveaxis(_to_u8(ref["rgb"][0]), -1, 0)
        assert np.sum(obs["camera_0"] != want) == 0
        frames.append(obs["camera_0"].astype(np.int32))
        msgs.append(env.draw_msg)
    changed = np.abs(frames[0] - frames[1]).sum(0) > 30
    assert changed.sum() > 40                                       # the block moved in the image
    red = (frames[0][0] > 180) & (frames[0][1] < 150) & (frames[0][2] < 150)
    assert red.sum() > 20                                           # and it is the reference's mesh colour (0.956, 0.396, 0.365)
    vec = VecSplatRenderer(env.splat_handler, cams, n_envs=2, fov=env.ch.camera.fov, background=env.splat_handler.server.background_rgb)
    out = vec.render(msgs)["camera_0"]
    for e in range(2):
        assert np.abs(out[e].astype(np.int32) - frames[e]).max() <= 1   # device fp64 composition vs host quaternion round trip
    # the same step from pre-batched body poses [E,L,7] (no draw messages, no per-env Python): host array and CUDA tensor
    bodies = np.stack([env.splat_handler.body_poses_from_msg(m) for m in msgs])
    assert np.array_equal(vec.render(bodies)["camera_0"], out)
    assert np.array_equal(vec.render(torch.from_numpy(bodies).cuda())["camera_0"], out)
    vec.close()
    env.close()


def test_headless_client_renders_an_injected_mesh():
    """The viser surface alone (INTEGRATION.md option 1): splats + a trimesh-like mesh node, client.get_render."""
    from sim_a_splat_b200.headless_viser import HeadlessViser

    class Quad:
        vertices = np.array([[-0.1, -0.1, 0], [0.1, -0.1, 0], [0.1, 0.1, 0], [-0.1, 0.1, 0]], dtype=np.float64)
        faces = np.array([[0, 1, 2], [0, 2, 3]])
        visual = type("V", (), {"vertex_colors": np.tile(np.uint8([30, 200, 40, 255]), (4, 1))})()

    rng = np.random.default_rng(1)
    srv = HeadlessViser(background_rgb=(0.0, 0.0, 0.0))
    srv.scene.add_gaussian_splats("/bg", centers=rng.uniform(-0.5, 0.5, size=(400, 3)) + [0, 0, 1.5], covariances=np.tile(np.eye(3) * 4e-4, (400, 1, 1)),
                                  rgbs=rng.uniform(0.2, 0.9, size=(400, 3)), opacities=np.full((400, 1), 0.6))
    h = srv.scene.add_mesh_trimesh("/mesh", Quad(), scale=1.0, position=(0.0, 0.0, 1.0))
    ch = srv.get_clients()[0]
    img0 = ch.get_render(height=96, width=128, wxyz=np.array([1.0, 0, 0, 0]), position=np.zeros(3))
    assert img0.shape == (96, 128, 3) and img0.dtype == np.uint8
    green = (img0[..., 1] > 150) & (img0[..., 0] < 90) & (img0[..., 2] < 100)
    assert green.sum() > 100 and green[48, 64]                    # the quad faces the camera at the image centre
    h.position = (0.25, 0.0, 1.0)                                  # node transform moves it to the right (+x)
    img1 = ch.get_render(height=96, width=128, wxyz=np.array([1.0, 0, 0, 0]), position=np.zeros(3))
    g1 = (img1[..., 1] > 150) & (img1[..., 0] < 90) & (img1[..., 2] < 100)
    assert g1.sum() > 100 and not g1[48, 64] and np.nonzero(g1)[1].mean() > np.nonzero(green)[1].mean() + 10
    ref = oracle.render(srv.splat_scene.as_oracle_dict(), np.eye(4, dtype=np.float32)[None], _K(96, 128, ch.camera.fov)[None], 128, 96,
                        srv.link_poses(), [0], bg=(0.0, 0.0, 0.0))
    assert np.sum(img1 != _to_u8(ref["rgb"][0])) == 0
    srv.stop()


def test_device_pose_composition_matches_pose_oracle(toy, assets):
    """sas_set_link_poses_from_bodies (csrc/sas_pose.cu) == oracle/pose_oracle.py:link_transform (pinned to the
    reference's draw_handler by tests/golden/draw_handler.npz), over many envs, incl. un-normalised quaternions."""
    from oracle import pose_oracle as po
    from sim_a_splat_b200.renderer import SplatRenderer
    from sim_a_splat_b200.splat.splat_handler import RigidTransformLite, SplatHandler

    h = SplatHandler(toy["root"], toy["obj"], toy["cfg"], toy["package_path"], toy["package_name"], toy["urdf_name"],
                     task_assets_name="tblock_paper.obj", sim_robot_weld_frame_transform=RigidTransformLite(translation=(0.03, 0.01, -0.02)))
    sc = h.server.build_scene()
    E = 37
    r = SplatRenderer(sc, max_images=E, max_envs=E, max_width=64, max_height=64)
    m = h.pose_model()
    r.set_pose_model(**m)
    rng = np.random.default_rng(5)
    body = np.concatenate([rng.normal(size=(E, N_LINKS, 4)) * rng.uniform(0.2, 3.0, size=(E, N_LINKS, 1)), rng.uniform(-0.6, 0.6, size=(E, N_LINKS, 3))], axis=-1)
    r.set_link_poses_from_bodies(body)
    got = r.link_poses(E)
    want = np.empty_like(got)
    for e in range(E):
        for l in range(N_LINKS):
            rot, pos = po.link_transform(body[e, l, :4], body[e, l, 4:], m["body_offset"][l], m["scale"], m["Ri"], m["ti"], m["fk_R"][l], m["fk_t"][l])
            want[e, l] = po.poses_to_rot12(rot, pos)
    np.testing.assert_allclose(got, want, atol=3e-7, rtol=0)
    # device-resident input gives the same block
    r.set_link_poses_from_bodies(torch.from_numpy(body).cuda())
    np.testing.assert_array_equal(r.link_poses(E), got)
    with pytest.raises(Exception, match="max_envs"):
        r.set_link_poses_from_bodies(np.zeros((E + 1, N_LINKS, 7)))
    r.close()


def test_unproject_depth_is_the_reference_pinhole_formula():
    """nerfstudio_utils.py:441-452: x=(u-cx)d/fx, y=(v-cy)d/fy, z=d; mask d<max_depth (bit-exact in fp32)."""
    from sim_a_splat_b200.renderer import SplatRenderer

    sc = S.make_synthetic_scene(2000, 1, seed=3)
    H, W = 37, 50
    K = np.float32([[61.5, 0, 24.25], [0, 60.25, 18.5], [0, 0, 1]])
    r = SplatRenderer(sc, max_images=2, max_width=W, max_height=H)
    r.set_cameras(S.ring_cameras(2, (0, 0, 0), seed=1), K, W, H)
    depth = r.render(("depth",))["depth"]
    xyz, valid = r.unproject_depth(depth, K, max_depth=1.1)
    d = depth.cpu().numpy()
    u, v = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    want = np.stack([(u - K[0, 2]) * d / K[0, 0], (v - K[1, 2]) * d / K[1, 1], d], axis=-1).astype(np.float32)
    np.testing.assert_array_equal(xyz.cpu().numpy(), want)
    np.testing.assert_array_equal(valid.cpu().numpy(), d < np.float32(1.1))
    assert 0 < int(valid.sum()) < valid.numel()
    r.close()


def test_gaussian_splat_render_contract(toy, assets):
    """GaussianSplat.render(pose) (nerfstudio_utils.py:123-177): OpenGL c2w in, splatfacto output dict out."""
    from pathlib import Path

    H, W = 90, 120
    K = _K(H, W, np.radians(60.0))
    c2w_cv = np.linalg.inv(S.lookat_viewmat((1.6, 0.4, 0.5), (0, 0, 0)).astype(np.float64))
    c2w_gl = c2w_cv.copy()
    c2w_gl[:3, 1:3] *= -1.0  # OpenCV -> OpenGL camera axes (y up, z backwards)
    gs = GaussianSplat(Path(toy["root"]) / "splatfacto" / toy["cfg"], intrinsics=(H, W, K), camera_to_worlds=c2w_gl[None, :3])
    h, w, Kt = gs.get_camera_intrinsics()
    assert (h, w) == (H, W) and Kt.shape == (3, 3) and gs.get_poses().shape == (1, 3, 4)
    out = gs.render(torch.from_numpy(c2w_gl[:3]).float())
    assert out["rgb"].shape == (H, W, 3) and out["depth"].shape == (H, W, 1) and out["accumulation"].shape == (H, W, 1)
    assert out["rgb"].is_cuda and out["background"].shape == (3,)
    sc = gs.scene
    vm = S.c2w_opengl_to_viewmat(c2w_gl[:3].astype(np.float32))[None]
    np.testing.assert_allclose(vm[0], np.linalg.inv(c2w_cv), atol=1e-5)
    ref = oracle.render(sc.as_oracle_dict(), vm, K[None], W, H)
    ns = oracle.oracle.nerfstudio_outputs(ref, NS_BACKGROUND)
    assert np.abs(out["rgb"].cpu().numpy() - ns["rgb"][0]).max() <= 1e-3
    assert np.abs(out["depth"].cpu().numpy() - ns["depth"][0]).max() <= 1e-3
    np.testing.assert_array_equal(out["accumulation"].cpu().numpy()[..., 0], ref["alpha"][0])
    assert float(out["rgb"].min()) >= 0.0 and float(out["rgb"].max()) <= 1.0
    # depth convention of generate_RGBD_point_cloud (:441-452): z along the optical axis -> unprojected points
    # of opaque pixels lie in front of the camera within the scene bounds
    d = out["depth"].cpu().numpy()[..., 0]
    a = ref["alpha"][0]
    assert np.all(d[a > 0.99] > 0.3) and np.all(d[a > 0.99] < 4.0)
    # res_factor rescales intrinsics and size (Cameras.rescale_output_resolution, :99-100)
    half = GaussianSplat(sc, res_factor=0.5, intrinsics=(H, W, K))
    assert half.get_camera_intrinsics()[:2] == (H // 2, W // 2)
    o2 = half.render(c2w_gl[:3])
    assert o2["rgb"].shape == (H // 2, W // 2, 3)
    pts, cols, dimg, rgbimg = gs.generate_RGBD_point_cloud(torch.from_numpy(c2w_gl[:3]).float(), max_depth=1.5)
    n_valid = int((d < 1.5).sum())
    assert pts.shape == (n_valid, 3) and cols.shape == (n_valid, 3) and 0 < n_valid < H * W
    assert float(pts[:, 2].max()) < 1.5 and torch.equal(dimg, out["depth"][..., 0])


def test_device_urdf_fk_matches_host_body_pose_source(toy, assets):
    """SURVEY 8f-1 / VERDICT r1 missing #5: joint angles [E,J] -> body poses -> link poses entirely on the device
    (csrc/sas_pose.cu fk_kernel + link_pose_kernel) == the host chain urdf_fk.BatchedBodyPoseSource.body_poses ->
    sas_set_link_poses_from_bodies, incl. the prismatic joint, the fixed joint and a free body (the T-block)."""
    from sim_a_splat_b200.renderer import SplatRenderer
    from sim_a_splat_b200.splat.splat_handler import RigidTransformLite, SplatHandler
    from sim_a_splat_b200.urdf_fk import BatchedBodyPoseSource, SimpleURDF

    h = SplatHandler(toy["root"], toy["obj"], toy["cfg"], toy["package_path"], toy["package_name"], toy["urdf_name"],
                     task_assets_name="tblock_paper.obj", sim_robot_weld_frame_transform=RigidTransformLite(translation=(0.03, 0.01, -0.02)))
    sc = h.server.build_scene()
    u = SimpleURDF.load(f"{toy['package_path']}/{toy['package_name']}urdf/{toy['urdf_name']}")
    base = np.eye(4); base[:3, 3] = (0.05, -0.02, 0.01)
    src = BatchedBodyPoseSource(u, base=base)
    E, L = 41, sc.L
    n_robot = min(L, len(src.bodies))
    if n_robot == L:
        n_robot = L - 1          # keep one free body in the test: the last link group is then posed by free_bodies
    r = SplatRenderer(sc, max_images=E, max_envs=E, max_width=64, max_height=64)
    r.set_pose_model(**h.pose_model())
    r.set_fk_model(src, n_bodies=n_robot)
    rng = np.random.default_rng(17)
    q = rng.uniform(-2.5, 2.5, size=(E, len(src.joint_names))) * np.array([1, 1, 0.1])
    free = np.concatenate([rng.normal(size=(E, L - n_robot, 4)), rng.uniform(-0.5, 0.5, size=(E, L - n_robot, 3))], axis=-1)
    # host chain
    body = np.concatenate([src.body_poses(q, n_links=n_robot), free], axis=1)
    r.set_link_poses_from_bodies(body)
    want = r.link_poses(E).copy()
    # device chain, host inputs then device inputs
    r.set_link_poses_from_joints(q, free)
    got = r.link_poses(E).copy()
    np.testing.assert_allclose(got, want, atol=2e-7, rtol=0)
    r.set_link_poses_from_joints(torch.from_numpy(q).cuda(), torch.from_numpy(free).cuda())
    np.testing.assert_array_equal(r.link_poses(E), got)
    # free bodies default to the identity pose
    r.set_link_poses_from_joints(q)
    ident = np.zeros((E, L - n_robot, 7)); ident[..., 0] = 1.0
    r.set_link_poses_from_bodies(np.concatenate([src.body_poses(q, n_links=n_robot), ident], axis=1))
    want_i = r.link_poses(E).copy()
    r.set_link_poses_from_joints(q)
    np.testing.assert_allclose(r.link_poses(E), want_i, atol=2e-7, rtol=0)
    r.close()
