"""HeadlessViser: the viser surface the reference touches, backed by the B200 renderer instead of a
browser tab.

The reference drives its live render path through these calls only (all in /root/reference):
  server.scene.add_gaussian_splats(name, centers, covariances, rgbs, opacities[, wxyz])
                                                          sim_a_splat/splat/splat_handler.py:106-119, :134-141
  handle.wxyz = ..., handle.position = ...                sim_a_splat/splat/splat_handler.py:241-249, :283-288
  server.scene.add_mesh_trimesh(name, mesh, scale)        sim_a_splat/splat/splat_handler.py:190-194, :215-219
  server.get_clients() -> {id: client}                    sim_a_splat/env/splat/splat_env_wrapper.py:88
  client.camera.position / .wxyz                          sim_a_splat/env/splat/splat_env_wrapper.py:57-64
  client.get_render(height, width, wxyz, position) -> uint8 [H,W,3]
                                                          sim_a_splat/env/splat/splat_env_wrapper.py:148-157
  server.stop()                                           sim_a_splat/env/splat/splat_env_wrapper.py:162

Semantics kept: a splat node registered with a `wxyz` argument is a rigid group whose node transform
(wxyz, position) moves its Gaussians (mu' = R mu + p, Sigma' = R Sigma R^T); nodes registered without
one are static background.  A triangle mesh registered with `add_mesh_trimesh` (the reference's robot
link meshes and T-block, splat_handler.py:190-194, :215-219) is drawn by the browser over the splat; here it
becomes a rigid group of surface Gaussians in its mean vertex colour (ellipsoids/mesh_utils.py), scaled by
`scale` and moved by the node transform -- pass `meshes_as_splats=False` to ignore meshes instead.  The viser
protocol carries covariances and DC colours only, so a node's
(quaternion, scale) are recovered from its covariances by a symmetric eigendecomposition and its SH is
degree 0 -- unless the caller passes the richer `quats= / scales= / sh=` keywords (our SplatHandler
does when the loader has them).  Triangle meshes are accepted and ignored (next).

viser details that the reference leaves implicit are explicit attributes here: the client's vertical
field of view (`client.camera.fov`, radians, viser default 75 deg [RECALLED]) and the canvas
background (`server.background_rgb`, default white [RECALLED]).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

from .scene import C0, SplatScene, c2w_opencv_to_viewmat
from .transforms import SO3


def covariances_to_quat_scale(covs: np.ndarray):
    """Sigma = R diag(s^2) R^T  ->  (wxyz, s): batched symmetric eigendecomposition, det(R) forced to +1."""
    covs = np.asarray(covs, dtype=np.float64).reshape(-1, 3, 3)
    w, v = np.linalg.eigh(0.5 * (covs + covs.transpose(0, 2, 1)))
    flip = np.linalg.det(v) < 0
    v[flip, :, 0] *= -1.0
    quats = np.stack([SO3.from_matrix(m).wxyz for m in v]) if len(v) else np.zeros((0, 4))
    return quats.astype(np.float32), np.sqrt(np.maximum(w, 0.0)).astype(np.float32)


class _NodeHandle:
    """What add_gaussian_splats / add_mesh_trimesh return: .wxyz / .position are the node transform."""

    def __init__(self, server: "HeadlessViser", name: str, wxyz=None, position=None):
        self._server, self.name = server, name
        self._wxyz = np.array([1.0, 0.0, 0.0, 0.0]) if wxyz is None else np.asarray(wxyz, dtype=np.float64)
        self._position = np.zeros(3) if position is None else np.asarray(position, dtype=np.float64)

    @property
    def wxyz(self):
        return self._wxyz

    @wxyz.setter
    def wxyz(self, v):
        self._wxyz = np.asarray(v, dtype=np.float64).reshape(4)

    @property
    def position(self):
        return self._position

    @position.setter
    def position(self, v):
        self._position = np.asarray(v, dtype=np.float64).reshape(3)

    def remove(self):
        self._server._remove(self.name)


class _SplatNode(_NodeHandle):
    def __init__(self, server, name, means, quats, scales, opacities, sh, dynamic, wxyz, position):
        super().__init__(server, name, wxyz, position)
        self.means, self.quats, self.scales, self.opacities, self.sh, self.dynamic = means, quats, scales, opacities, sh, dynamic


class _Scene:
    def __init__(self, server: "HeadlessViser"):
        self._s = server

    def add_gaussian_splats(self, name: str, centers, covariances=None, rgbs=None, opacities=None, wxyz=None, position=None,
                            quats=None, scales=None, sh=None):
        centers = np.asarray(centers, dtype=np.float32).reshape(-1, 3)
        n = centers.shape[0]
        if quats is None or scales is None:
            quats, scales = covariances_to_quat_scale(covariances)
        if sh is None:
            sh = ((np.asarray(rgbs, dtype=np.float64).reshape(n, 3) - 0.5) / C0).astype(np.float32)[:, None, :]
        node = _SplatNode(self._s, name, centers, np.asarray(quats, dtype=np.float32).reshape(n, 4),
                          np.asarray(scales, dtype=np.float32).reshape(n, 3), np.asarray(opacities, dtype=np.float32).reshape(n),
                          np.asarray(sh, dtype=np.float32).reshape(n, -1, 3), wxyz is not None or position is not None, wxyz, position)
        self._s._add(node)
        return node

    def add_mesh_trimesh(self, name: str, mesh=None, scale=1.0, wxyz=None, position=None):
        verts, faces = getattr(mesh, "vertices", None), getattr(mesh, "faces", None)
        if mesh is None or verts is None or faces is None or not self._s.meshes_as_splats:
            h = _NodeHandle(self._s, name, wxyz, position)  # pose bookkeeping only: nothing is drawn
            self._s.mesh_handles[name] = h
            return h
        from .ellipsoids.mesh_utils import TriMesh, mesh_to_gaussians

        rgb = (0.7, 0.7, 0.7)
        vc = getattr(getattr(mesh, "visual", None), "vertex_colors", None)
        if vc is not None and len(vc):
            vc = np.asarray(vc, dtype=np.float64)[:, :3]
            rgb = tuple((vc.mean(0) / (255.0 if vc.max() > 1.0 else 1.0)).tolist())
        tm = TriMesh(np.asarray(verts, dtype=np.float64) * float(scale), faces)
        a, b, c = tm.vertices[tm.faces[:, 0]], tm.vertices[tm.faces[:, 1]], tm.vertices[tm.faces[:, 2]]
        area = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1).sum()
        n = int(np.clip(area / (self._s.mesh_gaussian_spacing ** 2), 200, 20000))
        means, quats, scales, opac, sh = mesh_to_gaussians(tm, n, rgb=rgb, seed=len(self._s.nodes))
        node = _SplatNode(self._s, name, means, quats, scales, opac, sh, True,
                          np.array([1.0, 0.0, 0.0, 0.0]) if wxyz is None else wxyz, position)
        self._s._add(node)
        self._s.mesh_handles[name] = node
        return node


class _Camera:
    def __init__(self):
        self.position = np.zeros(3)
        self.wxyz = np.array([1.0, 0.0, 0.0, 0.0])
        self.fov = np.radians(75.0)


class _Client:
    def __init__(self, server: "HeadlessViser"):
        self._s, self.camera = server, _Camera()

    def get_render(self, height: int, width: int, wxyz=None, position=None, transport_format: str = "jpeg") -> np.ndarray:
        wxyz = self.camera.wxyz if wxyz is None else wxyz
        position = self.camera.position if position is None else position
        return self._s.render_views([(int(height), int(width), np.asarray(wxyz), np.asarray(position))], self.camera.fov)[0]


class HeadlessViser:
    """Stands where `viser.ViserServer()` stands (splat_handler.py:35-43)."""

    def __init__(self, device: int = 0, background_rgb=(1.0, 1.0, 1.0), max_isects: int = 0, meshes_as_splats: bool = True,
                 mesh_gaussian_spacing: float = 0.004):
        """mesh_gaussian_spacing: target distance between the surface Gaussians a mesh is turned into (scene units)."""
        self.scene = _Scene(self)
        self.meshes_as_splats, self.mesh_gaussian_spacing = bool(meshes_as_splats), float(mesh_gaussian_spacing)
        self.device, self.background_rgb, self._max_isects = device, tuple(float(c) for c in background_rgb), max_isects
        self.nodes: Dict[str, _SplatNode] = {}
        self.mesh_handles: Dict[str, _NodeHandle] = {}
        self._client = _Client(self)
        self._renderer = None
        self._dirty = True
        self._dynamic: List[_SplatNode] = []
        self.splat_scene: Optional[SplatScene] = None

    # -- viser surface ------------------------------------------------------------------------
    def get_clients(self):
        return {0: self._client}

    def stop(self):
        if self._renderer is not None:
            self._renderer.close()
            self._renderer = None

    # -- internals ----------------------------------------------------------------------------
    def _add(self, node: _SplatNode):
        self.nodes[node.name] = node
        self._dirty = True

    def _remove(self, name: str):
        self.nodes.pop(name, None)
        self.mesh_handles.pop(name, None)
        self._dirty = True

    def build_scene(self) -> SplatScene:
        """Canonical order: static nodes first (registration order), then one link per dynamic node."""
        static = [n for n in self.nodes.values() if not n.dynamic]
        self._dynamic = [n for n in self.nodes.values() if n.dynamic]
        order = static + self._dynamic
        if not order:
            raise RuntimeError("no Gaussian splats registered")
        K = max(n.sh.shape[1] for n in order)
        sh = [np.pad(n.sh, ((0, 0), (0, K - n.sh.shape[1]), (0, 0))) for n in order]
        link = [np.full(n.means.shape[0], -1 if not n.dynamic else self._dynamic.index(n), dtype=np.int32) for n in order]
        cat = lambda xs: np.concatenate(xs, axis=0)  # noqa: E731
        N = sum(n.means.shape[0] for n in order)
        self.splat_scene = SplatScene(cat([n.means for n in order]), cat([n.quats for n in order]), cat([n.scales for n in order]),
                                      cat([n.opacities for n in order]), cat(sh), cat(link), len(self._dynamic), np.arange(N))
        return self.splat_scene

    def link_poses(self) -> np.ndarray:
        """[1,L,12] fp32 from the dynamic nodes' current (wxyz, position)."""
        out = np.zeros((1, max(len(self._dynamic), 1), 12), dtype=np.float32)
        for l, n in enumerate(self._dynamic):
            out[0, l, :9] = SO3(n.wxyz).as_matrix().reshape(9)
            out[0, l, 9:] = n.position
        return out[:, : len(self._dynamic)]

    def render_views(self, views, fov: float, outputs=("rgb_u8",)):
        """views: [(H, W, wxyz, position)] camera-to-world, OpenCV axes.  Same-size views are rendered in one call."""
        from .renderer import SplatRenderer

        if self._dirty or self._renderer is None:
            if self._renderer is not None:
                self._renderer.close()
            sc = self.build_scene()
            self._renderer = SplatRenderer(sc, device=self.device, max_images=16, max_envs=1, max_width=4080, max_height=4080,
                                           max_isects=self._max_isects)
            self._dirty = False
        r = self._renderer
        if self._dynamic:
            r.set_link_poses(self.link_poses())
        imgs: List[Optional[np.ndarray]] = [None] * len(views)
        sizes = sorted({(v[0], v[1]) for v in views})
        for (H, W) in sizes:
            idx = [i for i, v in enumerate(views) if (v[0], v[1]) == (H, W)]
            fy = H / (2.0 * np.tan(fov / 2.0))
            Kmat = np.array([[fy, 0, W / 2.0], [0, fy, H / 2.0], [0, 0, 1.0]])
            vms = np.stack([c2w_opencv_to_viewmat(views[i][2], views[i][3]) for i in idx])
            for c0 in range(0, len(idx), 16):
                chunk = idx[c0:c0 + 16]
                r.set_cameras(vms[c0:c0 + 16], Kmat, W, H, np.zeros(len(chunk), dtype=np.int32))
                out = r.render(outputs, background=self.background_rgb, to_host=True)
                for k, i in enumerate(chunk):
                    imgs[i] = out[outputs[0]][k].numpy().copy() if len(outputs) == 1 else {o: out[o][k].numpy().copy() for o in outputs}
        return imgs
