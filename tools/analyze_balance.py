#!/usr/bin/env python
"""CPU work model of the compositor's intra-CTA balance: per 256-record batch the CTA waits for its slowest
warp.  Compares block->warp assignments.  (Test infrastructure; uses the oracle.)"""
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
    schemes = {
        "contig 8x4 warp, 2x2 groups": lambda cx, ry: ((ry >> 1) * 2 + (cx >> 1), (ry & 1) * 2 + (cx & 1)),
        "diagonal (ry+2cx)&7": lambda cx, ry: ((ry + 2 * cx) & 7, cx),
        "column stripes w=cx+4*(ry&1)": lambda cx, ry: (cx + 4 * (ry & 1), ry >> 1),
        "rows interleaved w=ry": lambda cx, ry: (ry, cx),
    }
    res = {k: [0, 0] for k in schemes}   # sum of max-over-warps, sum of all warps
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
            nb = (n + 255) // 256
            # per block (cx, ry): hits per 32-chunk [nchunks], zero after the block's stop, zero if static
            nch = nb * 8
            blk = np.zeros((4, 8, nch), dtype=np.int32)
            for cx in range(4):
                for ry in range(8):
                    x0, y0 = tx * 16 + 4 * cx, ty * 16 + 2 * ry
                    h = ((gx + gex >= x0 + 0.5) & (gx - gex <= x0 + 3.5) & (gy + gey >= y0 + 0.5) & (gy - gey <= y0 + 1.5))
                    if not (h & is_link).any():
                        continue  # static group
                    stop = int(stop_pix[2 * ry:2 * ry + 2, 4 * cx:4 * cx + 4].max())
                    h = h.copy(); h[stop:] = False
                    hp = np.zeros(nch * 32, dtype=np.int32); hp[:n] = h
                    blk[cx, ry] = hp.reshape(nch, 32).sum(1)
            # dynamic blocks packed densely into warps (k-th dynamic block in sas_group_bit order -> slot k)
            def bit_of(cx, ry):
                return ((ry >> 1) * 2 + (cx >> 1)) * 4 + (ry & 1) * 2 + (cx & 1)
            dyn_bits = sorted(bit_of(cx, ry) for cx in range(4) for ry in range(8) if blk[cx, ry].any())
            slot_of = {b: k for k, b in enumerate(dyn_bits)}
            nxt = [len(dyn_bits)]
            def packed(cx, ry):
                b = bit_of(cx, ry)
                if b in slot_of:
                    k = slot_of[b]
                else:
                    k = nxt[0]; nxt[0] += 1
                return (k >> 2, k & 3)
            for name, fn in list(schemes.items()) + [("packed dynamic blocks", packed)]:
                nxt[0] = len(dyn_bits)
                res.setdefault(name, [0, 0])
                wcost = np.zeros((8, 4, nch), dtype=np.int32)
                for cx in range(4):
                    for ry in range(8):
                        w, g = fn(cx, ry)
                        wcost[w, g] += blk[cx, ry]
                per_warp_chunk = wcost.max(1)                       # chunk-synchronous groups
                per_warp_batch = per_warp_chunk.reshape(8, nb, 8).sum(2)
                warp_has_dyn = wcost.sum((1, 2)) > 0
                per_warp_batch = per_warp_batch + (8 * 1) * warp_has_dyn[:, None]   # test phase ~ 1 walk-equivalent per chunk, skipped by idle warps
                res[name][0] += int(per_warp_batch.max(0).sum())
                res[name][1] += int(per_warp_batch.sum())
    for name, (mx, tot) in res.items():
        print(f"{name:34s} sum_batches max_warp = {mx:9d}   total work = {tot:9d}   balance = {tot / (8 * mx):.3f}")

if __name__ == "__main__":
    main()
