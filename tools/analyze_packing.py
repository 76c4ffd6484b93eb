#!/usr/bin/env python
"""CPU work model: does the ORDER in which a tile's link-reachable 4x2 groups are packed four per warp matter?
The four groups of a warp walk chunk-synchronously, so a warp's walk iterations per 32-record chunk are the longest
of its four hit lists.  Compares sas_group_bit order (what the kernel does) against orders by hit density.
(Test infrastructure; uses the oracle.)   python tools/analyze_packing.py [n_envs]"""
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
        m2 = f.means2d[i]; con = f.conics[i]; rad = f.radii[i]; op = sc.opacities
        A, B, Cc = con[:, 0].astype(np.float64), con[:, 1].astype(np.float64), con[:, 2].astype(np.float64)
        detc = A * Cc - B * B
        with np.errstate(all="ignore"):
            a = Cc / detc; c = A / detc
            tau = np.maximum(np.log(255.0 * op) + 0.01, 0.0)
            ex = np.minimum(np.sqrt(2 * tau * a) * 1.001 + 0.02, rad[:, 0] + 1.0)
            ey = np.minimum(np.sqrt(2 * tau * c) * 1.001 + 0.02, rad[:, 1] + 1.0)
        for t in range(T):
            ids = f.flatten_ids[offs[i * T + t]:offs[i * T + t + 1]] - i * N
            is_link = ids >= nbg
            if ids.size == 0 or not is_link.any():
                continue
            ty, tx = divmod(t, tw)
            n = ids.size
            gx, gy, gex, gey = m2[ids, 0], m2[ids, 1], ex[ids], ey[ids]
            alpha = f.alpha[i, ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16]
            last = f.last_idx[i, ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16] - offs[i * T + t]
            stop_pix = np.clip(np.where(alpha >= 1 - 1.1e-4, last + 2, n), 0, n)
            nch = (n + 31) // 32
            groups = []   # (bit, hits per chunk incl. early stop, background hits without stop = what a per-view pass could know)
            for cx in range(4):
                for ry in range(8):
                    x0, y0 = tx * 16 + 4 * cx, ty * 16 + 2 * ry
                    h = ((gx + gex >= x0 + 0.5) & (gx - gex <= x0 + 3.5) & (gy + gey >= y0 + 0.5) & (gy - gey <= y0 + 1.5))
                    if not (h & is_link).any():
                        continue
                    bg_total = int((h & ~is_link).sum())
                    stop = int(stop_pix[2 * ry:2 * ry + 2, 4 * cx:4 * cx + 4].max())
                    h = h.copy(); h[stop:] = False
                    hp = np.zeros(nch * 32, dtype=np.int32); hp[:n] = h
                    bit = ((ry >> 1) * 2 + (cx >> 1)) * 4 + (ry & 1) * 2 + (cx & 1)
                    groups.append((bit, hp.reshape(nch, 32).sum(1), bg_total, cx, ry))
            if not groups:
                continue
            orders = {
                "sas_group_bit order (kernel)": sorted(range(len(groups)), key=lambda k: groups[k][0]),
                "row major (ry, cx)": sorted(range(len(groups)), key=lambda k: (groups[k][4], groups[k][3])),
                "column major (cx, ry)": sorted(range(len(groups)), key=lambda k: (groups[k][3], groups[k][4])),
                "8x8-pixel blocks (cx>>1, ry>>2), then inside": sorted(range(len(groups)), key=lambda k: (groups[k][4] >> 2, groups[k][3] >> 1, groups[k][4] & 3, groups[k][3] & 1)),
                "boustrophedon over 8x4-pixel blocks": sorted(range(len(groups)), key=lambda k: ((groups[k][4] >> 1), (groups[k][3] >> 1) if ((groups[k][4] >> 1) & 1) == 0 else 1 - (groups[k][3] >> 1), groups[k][4] & 1, groups[k][3] & 1)),
                "by view background hits, descending": sorted(range(len(groups)), key=lambda k: -groups[k][2]),
                "by actual hits (oracle knowledge)": sorted(range(len(groups)), key=lambda k: -int(groups[k][1].sum())),
            }
            for name, od in orders.items():
                r = res.setdefault(name, [0, 0, 0])
                ng = len(od)
                n_cta = 1 if ng <= 16 else 2
                per_cta = []
                for p in range(n_cta):
                    sl = od[p * 16:(p + 1) * 16]
                    warps = [sl[w * 4:(w + 1) * 4] for w in range(4) if sl[w * 4:(w + 1) * 4]]
                    wc = np.stack([np.max(np.stack([groups[k][1] for k in wg]), 0) for wg in warps])   # [warps, chunks]
                    r[1] += int(wc.sum())                                   # walk iterations (issue)
                    nb = (nch + 3) // 4                                     # 128-record batches
                    pad = np.zeros((wc.shape[0], nb * 4), np.int64); pad[:, :nch] = wc
                    per_cta.append(int(pad.reshape(wc.shape[0], nb, 4).sum(2).max(0).sum()))   # barrier per batch: slowest warp
                r[0] += sum(per_cta)
                r[2] += int(sum(int(g[1].sum()) for g in groups))
    base = None
    for name, (mk, it, hits) in res.items():
        base = base or (mk, it)
        print(f"{name:40s} walk iterations {it:10d} ({it / base[1]:.3f})  sum of per-batch slowest warp {mk:10d} ({mk / base[0]:.3f})  lane fill {hits / (4 * it):.3f}")


if __name__ == "__main__":
    main()
