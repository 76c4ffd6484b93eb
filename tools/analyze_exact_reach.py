#!/usr/bin/env python
"""CPU work model: reach masks of the BACKGROUND entries refined by the exact sigma test (computed once per view entry,
off the compositor's critical path) -- walk iterations and staged records of the packed compositor against the culling-box
masks the kernel uses today.  (Test infrastructure; uses the oracle.)   python tools/analyze_exact_reach.py [n_envs]"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from sim_a_splat_b200 import scene as S  # noqa: E402


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    W = H = 96
    sc = S.make_reference_like_scene("robots-scene-v2", 3, 102, with_tblock=True)
    cen = sc.means[sc.link_id >= 0].mean(0)
    cam = S.ring_cameras(1, cen, 8); K = S.fov_intrinsics(W, H)
    poses = S.random_link_poses(E, sc.L, seed=1000)
    f = oracle.render(sc.as_oracle_dict(), np.repeat(cam, E, 0), np.repeat(K[None], E, 0), W, H, poses, np.arange(E))
    N, nbg = sc.N, sc.n_background
    tw = f.tile_w; T = tw * f.tile_h
    offs = np.append(f.isect_offsets.ravel(), f.M)
    res = {}
    for i in range(E):
        m2 = f.means2d[i]; con = f.conics[i].astype(np.float64); rad = f.radii[i]; op = sc.opacities.astype(np.float64)
        A, B, Cc = con[:, 0], con[:, 1], con[:, 2]
        detc = A * Cc - B * B
        with np.errstate(all="ignore"):
            a = Cc / detc; c = A / detc
            tau = np.maximum(np.log(255.0 * op) + 0.01, 0.0)
            ex = np.minimum(np.sqrt(2 * tau * a) * 1.001 + 0.02, rad[:, 0] + 1.0)
            ey = np.minimum(np.sqrt(2 * tau * c) * 1.001 + 0.02, rad[:, 1] + 1.0)
            thr = np.minimum(5.55, np.log(255.0 * op) + 0.01)
        for t in range(T):
            ids = f.flatten_ids[offs[i * T + t]:offs[i * T + t + 1]] - i * N
            is_link = ids >= nbg
            if ids.size == 0 or not is_link.any():
                continue
            ty, tx = divmod(t, tw)
            n = ids.size
            gx, gy, gex, gey = m2[ids, 0].astype(np.float64), m2[ids, 1].astype(np.float64), ex[ids], ey[ids]
            alpha = f.alpha[i, ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16]
            last = f.last_idx[i, ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16] - offs[i * T + t]
            stop_pix = np.clip(np.where(alpha >= 1 - 1.1e-4, last + 2, n), 0, n)
            nch = (n + 31) // 32
            box = {}; exact = {}
            for cx in range(4):
                for ry in range(8):
                    x0, y0 = tx * 16 + 4 * cx, ty * 16 + 2 * ry
                    h = ((gx + gex >= x0 + 0.5) & (gx - gex <= x0 + 3.5) & (gy + gey >= y0 + 0.5) & (gy - gey <= y0 + 1.5))
                    bit = ((ry >> 1) * 2 + (cx >> 1)) * 4 + (ry & 1) * 2 + (cx & 1)
                    box[bit] = h
                    he = np.zeros(n, bool)
                    for dy in range(2):
                        for dx in range(4):
                            ddx = gx - (x0 + dx + 0.5); ddy = gy - (y0 + dy + 0.5)
                            sig = 0.5 * (A[ids] * ddx * ddx + Cc[ids] * ddy * ddy) + B[ids] * ddx * ddy
                            he |= (sig >= 0) & (sig <= thr[ids])
                    exact[bit] = h & he
            dyn_bits = sorted(b for b in box if (box[b] & is_link).any())
            if not dyn_bits:
                continue
            variants = {
                "culling boxes (kernel today)": {b: box[b] for b in dyn_bits},
                "exact sigma test for background entries": {b: np.where(is_link, box[b], exact[b]) for b in dyn_bits},
                "exact sigma test for all entries": {b: exact[b] for b in dyn_bits},
            }
            for name, masks in variants.items():
                r = res.setdefault(name, [0, 0, 0])
                keep = np.zeros(n, bool)
                for b in dyn_bits:
                    keep |= masks[b]
                keep |= is_link if "all" not in name else keep   # link entries are never dropped by the merge
                r[2] += int(keep.sum()) * (1 if len(dyn_bits) <= 16 else 2)
                # compacted list: positions renumbered
                idx = np.nonzero(keep)[0]
                nk = idx.size; nchk = (nk + 31) // 32
                per_group = []
                for b in dyn_bits:
                    ry = (b >> 3) * 2 + ((b >> 1) & 1); cx = ((b >> 2) & 1) * 2 + (b & 1)
                    stop = int(stop_pix[2 * ry:2 * ry + 2, 4 * cx:4 * cx + 4].max())
                    h = masks[b].copy(); h[stop:] = False
                    hp = np.zeros(nchk * 32, np.int32); hp[:nk] = h[idx]
                    per_group.append(hp.reshape(nchk, 32).sum(1))
                for p in range(0, len(per_group), 4):
                    r[0] += int(np.max(np.stack(per_group[p:p + 4]), 0).sum())
                r[1] += int(sum(int(g.sum()) for g in per_group))
    base = None
    for name, (it, hits, staged) in res.items():
        base = base or (it, hits, staged)
        print(f"{name:44s} walk iterations {it:9d} ({it / base[0]:.3f})  hits {hits:9d} ({hits / base[1]:.3f})  staged records {staged:9d} ({staged / base[2]:.3f})")


if __name__ == "__main__":
    main()
