#!/usr/bin/env python
"""CPU work model of the compositor (K6) on BASELINE configs[1] -- how many warp-level "walk" iterations
different footprint partitions and the sub-tile static shortcut would cost.  Uses the oracle's
intermediates (test infrastructure); prints a table.  Not part of the product path.

    python tools/analyze_composite.py [n_envs]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from sim_a_splat_b200 import scene as S  # noqa: E402


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    W = H = 96
    sc = S.make_reference_like_scene("robots-scene-v2", 3, 102, with_tblock=True)
    cen = sc.means[sc.link_id >= 0].mean(0)
    cam = S.ring_cameras(1, cen, 8)
    K = S.fov_intrinsics(W, H)
    poses = S.random_link_poses(E, sc.L, seed=1000)
    f = oracle.render(sc.as_oracle_dict(), np.repeat(cam, E, 0), np.repeat(K[None], E, 0), W, H, poses, np.arange(E))
    N, nbg = sc.N, sc.n_background
    tw, th = f.tile_w, f.tile_h
    T = tw * th
    offs = np.append(f.isect_offsets.ravel(), f.M)
    tot = dict(base=0, static_warp=0, half=0, quarter=0, quarter_static=0, pix=0, tests=0, lists=0, n_dyn_tiles=0, n_static_tiles=0,
               warps=0, warps_static=0, base_dyn=0, tight=0, tight_quarter_static=0)
    for i in range(E):
        m2 = f.means2d[i]; con = f.conics[i]; rad = f.radii[i]
        op = sc.opacities
        # extents as in sas_project.cu (a, c = blurred cov diag, recover from conic: a = conic_c/det', ...)
        A, B, Cc = con[:, 0].astype(np.float64), con[:, 1].astype(np.float64), con[:, 2].astype(np.float64)
        detc = A * Cc - B * B
        with np.errstate(all="ignore"):
            a = Cc / detc; c = A / detc
            tau = np.maximum(np.log(255.0 * op) + 0.01, 0.0)
            ex = np.minimum(np.sqrt(2 * tau * a) * 1.001 + 0.02, rad[:, 0] + 1.0)
            ey = np.minimum(np.sqrt(2 * tau * c) * 1.001 + 0.02, rad[:, 1] + 1.0)
        for t in range(T):
            ids = f.flatten_ids[offs[i * T + t]:offs[i * T + t + 1]] - i * N
            if ids.size == 0:
                continue
            ty, tx = divmod(t, tw)
            is_link = ids >= nbg
            if not is_link.any():
                tot["n_static_tiles"] += 1
                continue
            tot["n_dyn_tiles"] += 1
            last = f.last_idx[i, ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16]  # per pixel stop index (abs into list?)
            n_cont = f.n_contrib[i, ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16]
            # per-pixel number of list entries visited: the oracle stops a pixel at its terminating Gaussian
            # approximate with: pixel processes the whole list unless saturated; use last_idx+1 if alpha high
            alpha = f.alpha[i, ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16]
            gx, gy = m2[ids, 0], m2[ids, 1]
            gex, gey = ex[ids], ey[ids]
            hA, gB, hC, gtau = 0.5 * A[ids], B[ids], 0.5 * Cc[ids], tau[ids]
            n = ids.size
            tot["lists"] += n
            # pixel stop positions: saturated pixels (T <= 1e-4 reached) stop right after last_idx; others go to n
            rel_last = last - offs[i * T + t]
            stop_pix = np.where(alpha >= 1 - 1.1e-4, rel_last + 2, n)
            stop_pix = np.clip(stop_pix, 0, n)
            # tight test data: per gaussian, sigma at nearest point is expensive; use exact ellipse bound per rect:
            for wy in range(4):
                for wx in range(2):
                    fx0, fy0 = tx * 16 + wx * 8, ty * 16 + wy * 4
                    stop_w = int(stop_pix[wy * 4:wy * 4 + 4, wx * 8:wx * 8 + 8].max())
                    sl = slice(0, stop_w)

                    def hits(x0, y0, w, h):
                        return ((gx[sl] + gex[sl] >= x0 + 0.5) & (gx[sl] - gex[sl] <= x0 + w - 0.5) &
                                (gy[sl] + gey[sl] >= y0 + 0.5) & (gy[sl] - gey[sl] <= y0 + h - 0.5))

                    def tight(x0, y0, w, h):
                        """exact: min of sigma over the rectangle of pixel centres <= tau (and AABB)"""
                        bx0, bx1, by0, by1 = x0 + 0.5, x0 + w - 0.5, y0 + 0.5, y0 + h - 0.5
                        mx, my = gx[sl], gy[sl]
                        a_, b_, c_, t_ = hA[sl], gB[sl], hC[sl], gtau[sl]
                        cxp, cyp = np.clip(mx, bx0, bx1), np.clip(my, by0, by1)
                        inside = (cxp == mx) & (cyp == my)
                        dx = mx - cxp  # fixed dx on facing x-edge (0 if inside x range)
                        dy = my - cyp
                        with np.errstate(all="ignore"):
                            # x-edge candidate: px = cxp, optimise py
                            dys = -b_ * dx / (2 * c_)
                            pys = np.clip(my - dys, by0, by1)
                            d2 = my - pys
                            s1 = a_ * dx * dx + b_ * dx * d2 + c_ * d2 * d2
                            # y-edge candidate: py = cyp, optimise px
                            dxs = -b_ * dy / (2 * a_)
                            pxs = np.clip(mx - dxs, bx0, bx1)
                            d1 = mx - pxs
                            s2 = a_ * d1 * d1 + b_ * d1 * dy + c_ * dy * dy
                        s1 = np.where(dx != 0, s1, np.inf)
                        s2 = np.where(dy != 0, s2, np.inf)
                        smin = np.where(inside, 0.0, np.minimum(s1, s2))
                        return hits(x0, y0, w, h) & (smin <= t_)

                    hw = hits(fx0, fy0, 8, 4)
                    tot["tight"] += int(tight(fx0, fy0, 8, 4).sum())
                    tq = [int(tight(fx0 + 4 * (k & 1), fy0 + 2 * (k >> 1), 4, 2).sum()) for k in range(4)]
                    tot["tight_quarter"] = tot.get("tight_quarter", 0) + max(tq)
                    th2 = [int(tight(fx0, fy0 + 2 * k, 8, 2).sum()) for k in range(2)]
                    tot["tight_half"] = tot.get("tight_half", 0) + max(th2)
                    tpm = 0
                    for py in range(4):
                        for px in range(8):
                            tpm = max(tpm, int(tight(fx0 + px, fy0 + py, 1, 1).sum()))
                    tw_ = tight(fx0, fy0, 8, 4)
                    # warp-level exact test AND per-group AABB
                    aq = max(int((tw_ & hits(fx0 + 4 * (k & 1), fy0 + 2 * (k >> 1), 4, 2)).sum()) for k in range(4))
                    a8 = max(int((tw_ & hits(fx0 + 2 * (k & 3), fy0 + 2 * (k >> 2), 2, 2)).sum()) for k in range(8))
                    t8 = max(int(tight(fx0 + 2 * (k & 3), fy0 + 2 * (k >> 2), 2, 2).sum()) for k in range(8))
                    def chunked(masks, cs=32):
                        m = np.stack(masks).astype(np.int32)  # [groups, n]
                        pad = (-m.shape[1]) % cs
                        m = np.pad(m, ((0, 0), (0, pad)))
                        return int(m.reshape(m.shape[0], -1, cs).sum(2).max(0).sum())
                    qm = [tw_ & hits(fx0 + 4 * (k & 1), fy0 + 2 * (k >> 1), 4, 2) for k in range(4)]
                    m8 = [tw_ & hits(fx0 + 2 * (k & 3), fy0 + 2 * (k >> 2), 2, 2) for k in range(8)]
                    for cs in (32, 256):
                        tot[f"chunk{cs}_wtight_aabb_quarter"] = tot.get(f"chunk{cs}_wtight_aabb_quarter", 0) + chunked(qm, cs)
                        tot[f"chunk{cs}_wtight_aabb_2x2"] = tot.get(f"chunk{cs}_wtight_aabb_2x2", 0) + chunked(m8, cs)
                        tot[f"chunk{cs}_aabb_quarter"] = tot.get(f"chunk{cs}_aabb_quarter", 0) + chunked([hits(fx0 + 4 * (k & 1), fy0 + 2 * (k >> 1), 4, 2) for k in range(4)], cs)
                    def hits_all(x0, y0, w, h):
                        return ((gx + gex >= x0 + 0.5) & (gx - gex <= x0 + w - 0.5) & (gy + gey >= y0 + 0.5) & (gy - gey <= y0 + h - 0.5))
                    for name, rects in (("q4x2", [(fx0 + 4 * (k & 1), fy0 + 2 * (k >> 1), 4, 2) for k in range(4)]),
                                        ("g2x2", [(fx0 + 2 * (k & 3), fy0 + 2 * (k >> 2), 2, 2) for k in range(8)]),
                                        ("g8x1", [(fx0, fy0 + k, 8, 1) for k in range(4)]),
                                        ("g4x1", [(fx0 + 4 * (k & 1), fy0 + (k >> 1), 4, 1) for k in range(8)])):
                        dyn = [bool((hits_all(*r) & is_link).any()) for r in rects]
                        ms = [hits(*r) for r, d in zip(rects, dyn) if d]
                        tot[f"N_{name}_groups"] = tot.get(f"N_{name}_groups", 0) + len(rects)
                        tot[f"N_{name}_static"] = tot.get(f"N_{name}_static", 0) + len(rects) - len(ms)
                        for cs in (32, 256):
                            tot[f"F_{name}_chunk{cs}_static"] = tot.get(f"F_{name}_chunk{cs}_static", 0) + (chunked(ms, cs) if ms else 0)
                    tot["wtight_aabb_quarter"] = tot.get("wtight_aabb_quarter", 0) + aq
                    tot["wtight_aabb_2x2"] = tot.get("wtight_aabb_2x2", 0) + a8
                    tot["tight_2x2"] = tot.get("tight_2x2", 0) + t8
                    tot["tight_pix"] = tot.get("tight_pix", 0) + tpm
                    nh = int(hw.sum())
                    tot["warps"] += 1
                    tot["base"] += nh
                    tot["tests"] += (stop_w + 31) // 32
                    link_hit = bool((hw & is_link[sl]).any())
                    # static test must look at ALL link gaussians of the tile (not only up to stop)
                    hw_all = ((gx + gex >= fx0 + 0.5) & (gx - gex <= fx0 + 7.5) & (gy + gey >= fy0 + 0.5) & (gy - gey <= fy0 + 3.5))
                    link_hit_all = bool((hw_all & is_link).any())
                    if not link_hit_all:
                        tot["warps_static"] += 1
                    else:
                        tot["static_warp"] += nh
                    # half-warp masks: two 8x2 footprints
                    h2 = [int(hits(fx0, fy0 + 2 * k, 8, 2).sum()) for k in range(2)]
                    tot["half"] += max(h2)
                    # quarter-warp masks: four 4x2 footprints (lane remap) -- 4 wide x 2 high
                    q = [int(hits(fx0 + 4 * (k & 1), fy0 + 2 * (k >> 1), 4, 2).sum()) for k in range(4)]
                    tot["quarter"] += max(q)
                    if link_hit_all:
                        tot["quarter_static"] += max(q)
                        tot["wtight_aabb_quarter_static"] = tot.get("wtight_aabb_quarter_static", 0) + aq
                        tot["wtight_aabb_2x2_static"] = tot.get("wtight_aabb_2x2_static", 0) + a8
                        tot["tight_2x2_static"] = tot.get("tight_2x2_static", 0) + t8
                        tot["tight_quarter_static"] += max(tq)
                    # per-pixel lower bound: max over pixels of own hits
                    pm = 0
                    for py in range(4):
                        for px in range(8):
                            pm = max(pm, int(hits(fx0 + px, fy0 + py, 1, 1).sum()))
                    tot["pix"] += pm
    print(f"envs={E} dyn tiles={tot['n_dyn_tiles']} static tiles={tot['n_static_tiles']} list entries in dyn tiles={tot['lists']}")
    print(f"warps in dyn tiles={tot['warps']} of which static (no link hit)={tot['warps_static']} ({100 * tot['warps_static'] / max(tot['warps'], 1):.1f} %)")
    b = tot["base"]
    for k in ("base", "static_warp", "half", "quarter", "quarter_static", "pix", "tight", "tight_half", "tight_quarter", "tight_quarter_static", "tight_pix", "wtight_aabb_quarter", "wtight_aabb_quarter_static", "wtight_aabb_2x2", "wtight_aabb_2x2_static", "tight_2x2", "tight_2x2_static", "chunk32_aabb_quarter", "chunk32_wtight_aabb_quarter", "chunk32_wtight_aabb_2x2", "chunk256_aabb_quarter", "chunk256_wtight_aabb_quarter", "chunk256_wtight_aabb_2x2"):
        print(f"{k:16s} walk iterations = {tot[k]:10d}  ({100 * tot[k] / b:5.1f} % of base)")
    for k in sorted(tot):
        if k.startswith("F_"):
            print(f"{k:28s} walk iterations = {tot[k]:10d}  ({100 * tot[k] / b:5.1f} % of base)")
        if k.startswith("N_") and k.endswith("static"):
            print(k, tot[k], "of", tot[k.replace("static", "groups")])
    print(f"test iterations (32 gaussians each, per warp) = {tot['tests']}")


if __name__ == "__main__":
    main()
