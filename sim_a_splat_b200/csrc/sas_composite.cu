// sas_composite.cu -- K6: per-tile front-to-back alpha compositing of RGB + depth + alpha
// (gsplat 1.5.2 rasterize_to_pixels, RGB+ED), and the splatfacto get_outputs post-processing.
//
// CTAs (256 threads, 32 registers, 8 resident per SM) pull 16x16 tiles from the work list of sas_work.cu through one
// atomic counter, heaviest first; a CTA takes SAS_COMP_ITEMS items (default 1; 0 = persistent CTAs, grid = SMs x 8)
// and exits, the grid holds enough CTAs for every possible item.  Two kinds of item: an image tile whose link bucket is
// non-empty is composited from its merged (background + link) list into the image; a view tile is
// the background-only list of a camera, composited once into the per-view buffers and afterwards
// copied (copy_static_tiles_kernel) into every image of that view whose link bucket for the tile
// is empty, and into the 4x2-pixel groups of the other tiles that no link Gaussian can reach
// (tile_dyn_mask, from the projection kernel) -- those pixels are bit-identical by construction.
//
// Within a tile every 8-lane group of a warp owns one 4x2 pixel block, one pixel per lane.  The blocks that link
// Gaussians can reach (all blocks of a view tile) are PACKED four per warp in sas_group_bit order and the tile is
// composited by one CTA of 4 warps (up to 16 such blocks) or two (more), each staging the list for itself: warps that
// would mostly wait at the batch barrier no longer hold warp slots, and there is one CTA shape (16 resident per SM).
// (SAS_VARIANT bit 9; without it a tile is one CTA of 8 warps, each owning a fixed 8x4 footprint.)  Measured:
// 1.207 -> 1.144 ms per step alone, 1.089 -> 1.010 ms with two renders in flight.  (Tried on the way: a 4-warp
// instance for tiles of <= 16 blocks next to an 8-warp instance for the rest, 1.202 / 1.050; plus a 6-warp instance,
// 1.314 / 1.084; CTAs of 2 warps and up to four per tile, 1.195 / 1.070.)
// The depth-sorted ids are consumed in batches of 256: every
// thread gathers one 48 B projected record into shared memory with three 16 B cp.async copies
// (double buffered, the id for the batch after next prefetched into a register), so global latency
// overlaps the arithmetic of the current batch; the thread that staged a record also computes, once
// for the whole CTA, the 32-bit mask of the tile's groups the record's culling extents reach.
// Inside a batch a warp takes 32 records at a time, ballots its four mask bits (one hit mask per
// group), and every group walks only its own hits -- in order, so front-to-back order is preserved.
// A lane stops when T*(1-alpha) <= 1e-4 (its pixel centre becomes NaN, which fails every later
// range test by itself); a group / warp skips once all its lanes stopped; the CTA leaves the tile
// when __syncthreads_count says all 256 pixels stopped.
//
// The reach test is conservative: the extents (sas_project.cu) bound the region where
// alpha >= 1/255 is possible -- a pixel with |dx| > ex has sigma >= dx^2/(2 cov_xx) > ln(255*opacity),
// i.e. alpha < 1/255, so the reference rule would have skipped it as well (the extents carry a
// margin far above the rounding of sigma).  Likewise sigma > SAS_SIGMA_SKIP skips the exp.  Neither
// changes any output bit.
//
// Not bandwidth bound.  55 % of the warp-stall samples are the per-batch barrier: the CTA waits for
// the warp over the densest 4x2 block, whose dependent sigma -> exp -> alpha -> T chain is the
// critical path; the walk loop is sensitive to its own instruction count (trimmed: no finished-flag
// test, clamp-free exp, one-compare range test), the test phase and the scheduler balance are not.
// DESIGN.md "What bounds the compositor" and "Experiments that did not pay" list what was measured.
#include "sas_common.cuh"

#include <cuda_fp16.h>

#include <cstdlib>

namespace {

constexpr int CB = 256;  // threads of the copy kernel = pixels of a tile

__device__ __forceinline__ unsigned char to_u8(float v) {
    v = fminf(fmaxf(v, 0.0f), 1.0f);
    return (unsigned char)(v * 255.0f + 0.5f);
}

struct Pixel {
    float pxf, pyf;            // pixel centre; pxf = NaN once the pixel has finished (every sigma then fails its
                               // range test by itself: no separate flag to test in the walk loop)
    float T, ar, ag, ab, ad;   // transmittance and accumulators
    int nc;
    __device__ __forceinline__ bool done() const { return p_isnan(pxf); }
    static __device__ __forceinline__ bool p_isnan(float v) { return v != v; }
};

// Blend staged record j into this lane's pixel (gsplat rasterize_to_pixels inner body).
// COUNT = false drops the per-pixel contributor count (only produced when n_contrib is requested).
template <bool COUNT>
__device__ __forceinline__ void blend_one(const SasRecord *sb, int j, Pixel &p) {
    const float4 A = sb[j].a;
    const float2 B = *reinterpret_cast<const float2 *>(&sb[j].b);  // conic c / 2, opacity
    const float dx = A.x - p.pxf, dy = A.y - p.pyf;
    const float sigma = fmaf(fmaf(A.z, dx, A.w * dy), dx, (B.x * dy) * dy);
    // 0 <= sigma <= SAS_SIGMA_SKIP as ONE unsigned compare of the bit pattern: negative values and NaN (a finished
    // pixel) have larger patterns; sigma is never -0.0 (its last term (c/2 dy) dy is >= +0)
    if (__float_as_uint(sigma) <= __float_as_uint(SAS_SIGMA_SKIP)) {
        const float al = fminf(SAS_ALPHA_MAX, B.y * sas_exp_neg_small(-sigma));
        if (al >= SAS_ALPHA_MIN) {
            const float nT = p.T * (1.0f - al);
            if (nT <= SAS_T_MIN) {
                // the pixel has finished.  (asm volatile keeps this a predicated one-instruction block: written as a plain
                // assignment the compiler forms a select that costs three moves on EVERY blend)
                asm volatile("mov.b32 %0, 0x7fc00000;" : "=f"(p.pxf));
            } else {
                const float4 Cc = sb[j].c;
                const float vis = al * p.T;
                p.ar = fmaf(Cc.x, vis, p.ar);
                p.ag = fmaf(Cc.y, vis, p.ag);
                p.ab = fmaf(Cc.z, vis, p.ab);
                p.ad = fmaf(Cc.w, vis, p.ad);
                p.T = nT;
                if (COUNT) p.nc++;
            }
        }
    }
}

// Composite the cnt records of one staged batch into this lane's pixel (whole warp calls this).
// X0, Y0: centre of the first pixel of the warp's 8x4 footprint.
//
// MODE 0: one hit mask per warp (the 8x4 footprint); every lane walks every hit.
// MODE 1: the footprint is split into four 4x2 blocks, one per 8-lane group (= one LDS.128
//   phase, so the four groups reading four different records costs no extra shared-memory
//   wavefronts); each group gets its own hit mask from four ballots and walks only its own hits, the
//   groups diverge inside the 32-record chunk and reconverge after it.  CPU work model
//   (tools/analyze_composite.py): 66 % of the per-warp walk iterations on the pusht batch; measured
//   18.1 M -> 10.1 M walk iterations per launch together with the static groups.
// MODE 2: as MODE 1, but the hit test is not repeated by every warp: the thread that staged a record
//   computed its 32-bit reach mask over all the tile's groups once (s_reach), a warp just ballots its
//   own four bits (profile of MODE 1: 60 instructions per 32-record chunk per warp in the test phase,
//   18 % of the kernel's instructions).
// PACKED (MODE 2 only): `warp` carries, instead of the warp index, the four reach-mask bit positions of this warp's
// groups (8 bits each): the groups were packed densely into warps, see composite_kernel.
template <int MODE, bool COUNT, bool PACKED>
__device__ __forceinline__ void process_batch(const SasRecord *sb, const unsigned *s_reach, int warp, int cnt, Pixel &p, float X0, float Y0) {
    const int lane = threadIdx.x & 31;
    const int grp = lane >> 3;
    for (int j0 = 0; j0 < cnt; j0 += 32) {
        const unsigned dm = __ballot_sync(0xffffffffu, p.done());
        if (dm == 0xffffffffu) break;
        unsigned mask;
        if (MODE == 2) {
            unsigned b00, b10, b01, b11;
            if (PACKED) {
                const unsigned rb = j0 + lane < cnt ? s_reach[j0 + lane] : 0u;
                b00 = __ballot_sync(0xffffffffu, (rb >> (warp & 31)) & 1u);
                b10 = __ballot_sync(0xffffffffu, (rb >> ((warp >> 8) & 31)) & 1u);
                b01 = __ballot_sync(0xffffffffu, (rb >> ((warp >> 16) & 31)) & 1u);
                b11 = __ballot_sync(0xffffffffu, (rb >> ((warp >> 24) & 31)) & 1u);
            } else {
                const unsigned bits = (j0 + lane < cnt ? s_reach[j0 + lane] : 0u) >> (warp * 4);
                b00 = __ballot_sync(0xffffffffu, bits & 1u);
                b10 = __ballot_sync(0xffffffffu, bits & 2u);
                b01 = __ballot_sync(0xffffffffu, bits & 4u);
                b11 = __ballot_sync(0xffffffffu, bits & 8u);
            }
            mask = (grp & 2) ? ((grp & 1) ? b11 : b01) : ((grp & 1) ? b10 : b00);
            if (((dm >> (grp * 8)) & 0xffu) == 0xffu) mask = 0u;  // every pixel of this group has finished
        } else if (MODE == 1) {
            bool hx0 = false, hx1 = false, hy0 = false, hy1 = false;
            if (j0 + lane < cnt) {
                const float4 A = sb[j0 + lane].a;
                const float4 B = sb[j0 + lane].b;
                const float xl = A.x - B.z, xh = A.x + B.z, yl = A.y - B.w, yh = A.y + B.w;
                hx0 = (xh >= X0) && (xl <= X0 + 3.0f);
                hx1 = (xh >= X0 + 4.0f) && (xl <= X0 + 7.0f);
                hy0 = (yh >= Y0) && (yl <= Y0 + 1.0f);
                hy1 = (yh >= Y0 + 2.0f) && (yl <= Y0 + 3.0f);
            }
            const unsigned b00 = __ballot_sync(0xffffffffu, hx0 && hy0);
            const unsigned b10 = __ballot_sync(0xffffffffu, hx1 && hy0);
            const unsigned b01 = __ballot_sync(0xffffffffu, hx0 && hy1);
            const unsigned b11 = __ballot_sync(0xffffffffu, hx1 && hy1);
            mask = (grp & 2) ? ((grp & 1) ? b11 : b01) : ((grp & 1) ? b10 : b00);
            if (((dm >> (grp * 8)) & 0xffu) == 0xffu) mask = 0u;  // every pixel of this group has finished
        } else {
            bool hit = false;
            if (j0 + lane < cnt) {
                const float4 A = sb[j0 + lane].a;
                const float4 B = sb[j0 + lane].b;
                hit = (A.x + B.z >= X0) && (A.x - B.z <= X0 + 7.0f) && (A.y + B.w >= Y0) && (A.y - B.w <= Y0 + 3.0f);
            }
            mask = __ballot_sync(0xffffffffu, hit);
        }
        const SasRecord *sb1 = sb + j0 - 1;  // __ffs is one-based
        while (mask) {  // the groups diverge here; no warp-synchronous operation inside
            const int j = __ffs(mask);
            mask &= mask - 1;
            blend_one<COUNT>(sb1, j, p);
        }
        __syncwarp();
    }
}

// ---- MODE 3: "split" compositor -- only the transmittance chain is serial ------------------------------------------
// sigma, exp and alpha of a (pixel, record) pair do not depend on the pixel's transmittance T; only
// T' = T (1 - alpha), the T' <= 1e-4 stop test and the accumulation do.  MODE 2 evaluates all of it inside the
// per-group walk (lane = pixel): 18 of 32 lanes active per instruction, 7 of 32 while blending, ~220 cycles per
// dependent walk iteration (profiles/r01e_ncu_composite.txt).  Here a staged batch is processed per warp in two
// phases over the warp's (record, group) HITS -- the records whose culling box reaches one of the warp's four
// packed 4x2 groups, queued in (group, record) order:
//   phase A (lane = hit): sigma, exp, alpha for the group's 8 pixels of one hit per lane -- 32 hits side by
//     side, no divergence, 8 independent chains per lane; alpha rows (0 where the pair fails the range or the
//     alpha >= 1/255 test) and the hit's colour go to a per-warp scratch in shared memory, and 8 ballots turn the
//     pass bits into one 32-hit mask PER PIXEL;
//   phase B (lane = pixel): every pixel walks only ITS OWN passing hits, in order: one LDS of alpha, one LDS.128 of
//     the colour, T' = T (1 - alpha), the stop test, four FMAs -- ~6 dependent operations per hit.
// Every pixel performs exactly the operations MODE 2 (and the oracle) performs for it, in the same order, so the
// images stay bit-identical.
struct __align__(16) SplitScratch {   // per warp
    float alpha[32][8];               // alpha of hit k (row) for the 8 pixels of its group; 0 = does not contribute
    float4 col[32];                   // r, g, b, depth of hit k
    uint4 hitmask[4];                 // [group]: 4 x 32-record chunk masks of the batch: records whose culling box reaches the group
};

// position of the r-th (0-based) set bit of w; requires r < popc(w)
__device__ __forceinline__ unsigned sas_nth_set_bit(unsigned w, unsigned r) {
    unsigned pos = 0u, c;
    c = __popc(w & 0xffffu); if (r >= c) { r -= c; pos = 16u; w >>= 16; }
    c = __popc(w & 0xffu);   if (r >= c) { r -= c; pos += 8u; w >>= 8; }
    c = __popc(w & 0xfu);    if (r >= c) { r -= c; pos += 4u; w >>= 4; }
    c = __popc(w & 0x3u);    if (r >= c) { r -= c; pos += 2u; w >>= 2; }
    c = w & 1u;              if (r >= c) pos += 1u;
    return pos;
}

__device__ __forceinline__ float sas_lds_f32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 sas_lds_f32x4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// an opaque copy of a 32-bit value: the compiler cannot rematerialise it from its inputs (it otherwise recomputes
// shared-memory addresses -- S2R included -- inside the walk loop when registers are tight)
__device__ __forceinline__ unsigned sas_pin(unsigned v) {
    unsigned r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

// a_addr: shared address of ws.alpha[0][this lane's pixel], c_addr: of ws.col[0] (both pinned by the caller)
template <bool COUNT>
__device__ __forceinline__ void process_batch_split(const SasRecord *sb, const unsigned *s_reach, unsigned gbits, int cnt, Pixel &p,
                                                    SplitScratch &ws, const float2 *s_org, unsigned a_addr, unsigned c_addr) {
    const int lane = threadIdx.x & 31, grp = lane >> 3, gl = lane & 7;
    const unsigned dm = __ballot_sync(0xffffffffu, p.done());
    if (dm == 0xffffffffu) return;
    // hit masks: bit l of word c of group g = record 32 c + l of the batch reaches group g (its culling box);
    // groups whose pixels have all finished drop out.  ng[g] = hits of group g in this batch.
    unsigned ng[4] = {0u, 0u, 0u, 0u};
    uint4 mym = make_uint4(0u, 0u, 0u, 0u);  // lane g < 4 collects the four chunk masks of group g
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const unsigned rb = (c * 32 + lane < cnt) ? s_reach[c * 32 + lane] : 0u;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            unsigned b = __ballot_sync(0xffffffffu, (rb >> ((gbits >> (8 * g)) & 31u)) & 1u);
            if (((dm >> (8 * g)) & 0xffu) == 0xffu) b = 0u;
            ng[g] += __popc(b);
            if (lane == g) { if (c == 0) mym.x = b; else if (c == 1) mym.y = b; else if (c == 2) mym.z = b; else mym.w = b; }
        }
    }
    const unsigned b1 = ng[0], b2 = b1 + ng[1], b3 = b2 + ng[2], h = b3 + ng[3];  // first hit index of groups 1..3, total
    if (h == 0) return;
    if (lane < 4) ws.hitmask[lane] = mym;
    // hits are numbered in (group, record) order and taken 32 per round; this lane's group owns hits [my_lo, my_hi)
    const unsigned my_lo = grp == 0 ? 0u : grp == 1 ? b1 : grp == 2 ? b2 : b3;
    const unsigned my_hi = grp == 0 ? b1 : grp == 1 ? b2 : grp == 2 ? b3 : h;
    __syncwarp();
    for (unsigned k0 = 0; k0 < h; k0 += 32) {
        // ---- phase A: lane = hit k ----
        const unsigned k = k0 + lane;
        unsigned pass8 = 0u;
        if (k < h) {
            // hit k -> (group g, rank r among g's hits) -> (chunk, r-th set bit) -> record of the batch
            const unsigned g = (k >= b1 ? 1u : 0u) + (k >= b2 ? 1u : 0u) + (k >= b3 ? 1u : 0u);
            unsigned r = k - (g == 0u ? 0u : g == 1u ? b1 : g == 2u ? b2 : b3);
            const uint4 wm = ws.hitmask[g];
            const unsigned p0 = __popc(wm.x), p1 = p0 + __popc(wm.y), p2 = p1 + __popc(wm.z);
            unsigned w = wm.x, jb = 0u, sub = 0u;
            if (r >= p0) { w = wm.y; jb = 32u; sub = p0; }
            if (r >= p1) { w = wm.z; jb = 64u; sub = p1; }
            if (r >= p2) { w = wm.w; jb = 96u; sub = p2; }
            const SasRecord *rc = sb + (jb + sas_nth_set_bit(w, r - sub));
            const float4 A = rc->a;
            const float2 B = *reinterpret_cast<const float2 *>(&rc->b);  // conic c / 2, opacity
            const float2 o = s_org[g];                                   // centre of the group's first pixel
            float dxv[4], al[8];
#pragma unroll
            for (int ix = 0; ix < 4; ix++) dxv[ix] = A.x - (o.x + (float)ix);
#pragma unroll
            for (int iy = 0; iy < 2; iy++) {
                const float dy = A.y - (o.y + (float)iy);
                const float t1 = A.w * dy, t2 = (B.x * dy) * dy;
#pragma unroll
                for (int ix = 0; ix < 4; ix++) {
                    const float sigma = fmaf(fmaf(A.z, dxv[ix], t1), dxv[ix], t2);
                    const float a0 = fminf(SAS_ALPHA_MAX, B.y * sas_exp_neg_small(-sigma));
                    const bool ok = (__float_as_uint(sigma) <= __float_as_uint(SAS_SIGMA_SKIP)) && (a0 >= SAS_ALPHA_MIN);
                    al[iy * 4 + ix] = ok ? a0 : 0.0f;
                    pass8 |= ok ? (1u << (iy * 4 + ix)) : 0u;
                }
            }
            float4 *row = reinterpret_cast<float4 *>(ws.alpha[lane]);
            row[0] = make_float4(al[0], al[1], al[2], al[3]);
            row[1] = make_float4(al[4], al[5], al[6], al[7]);
            if (pass8) ws.col[lane] = rc->c;
        }
        // one 32-hit mask per pixel: bit kk = hit k0 + kk passes at THIS lane's pixel
        unsigned mine = 0u;
#pragma unroll
        for (int pp = 0; pp < 8; pp++) {
            const unsigned b = __ballot_sync(0xffffffffu, (pass8 >> pp) & 1u);
            if (gl == pp) mine = b;
        }
        {   // ... restricted to the hits of this lane's group
            const int lo = max((int)my_lo - (int)k0, 0), hi = min((int)my_hi - (int)k0, 32);
            mine = hi > lo ? (mine & ((0xffffffffu >> (32 - (hi - lo))) << lo)) : 0u;
        }
        if (p.done()) mine = 0u;
        __syncwarp();  // alpha rows and colours are visible to the warp
        // ---- phase B: lane = pixel, only the transmittance chain ----
        while (mine) {
            const unsigned kk = (unsigned)__ffs(mine) - 1u;
            mine &= mine - 1u;
            const float a = sas_lds_f32(a_addr + kk * 32u);
            const float nT = p.T * (1.0f - a);
            if (nT <= SAS_T_MIN) {
                asm volatile("mov.b32 %0, 0x7fc00000;" : "=f"(p.pxf));  // finished
                mine = 0u;
            } else {
                const float4 Cc = sas_lds_f32x4(c_addr + kk * 16u);
                const float vis = a * p.T;
                p.ar = fmaf(Cc.x, vis, p.ar);
                p.ag = fmaf(Cc.y, vis, p.ag);
                p.ab = fmaf(Cc.z, vis, p.ab);
                p.ad = fmaf(Cc.w, vis, p.ad);
                p.T = nT;
                if (COUNT) p.nc++;
            }
        }
        __syncwarp();  // scratch is free for the next round
    }
}

__device__ __forceinline__ void write_pixel(const SasOutputs &o, size_t pi, float r, float g, float bl, float depth, float alpha, int nc) {
    if (o.rgb) { o.rgb[3 * pi + 0] = r; o.rgb[3 * pi + 1] = g; o.rgb[3 * pi + 2] = bl; }
    if (o.alpha) o.alpha[pi] = alpha;
    if (o.depth) o.depth[pi] = depth;
    if (o.rgb_u8) { o.rgb_u8[3 * pi + 0] = to_u8(r); o.rgb_u8[3 * pi + 1] = to_u8(g); o.rgb_u8[3 * pi + 2] = to_u8(bl); }
    if (o.n_contrib) o.n_contrib[pi] = nc;
    if (o.depth_f16) o.depth_f16[pi] = __half_as_ushort(__float2half_rn(depth));
}

// ---- bulk-async staging (SAS_VARIANT bit 11): one 48-byte cp.async.bulk per record, completing on an mbarrier with
// expect_tx, instead of three 16-byte cp.async (LDGSTS) + wait_group.  TMA *tensor* copies do not apply (the records
// of a tile list are a data-dependent gather), the 1-D bulk copy does: UBLKCP + SYNCS in SASS.
__device__ __forceinline__ void sas_mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void sas_mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sas_mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SAS_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra SAS_WAIT_%=;\n"
        "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void sas_bulk_copy48(void *smem_dst, const void *gmem_src, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 48, [%2];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

__device__ __forceinline__ void stage_record(SasRecord *dst, const SasRecord *src) {
    sas_cp_async16(&dst->a, &src->a);
    sas_cp_async16(&dst->b, &src->b);
    sas_cp_async16(&dst->c, &src->c);
}

// WARPS = warps per CTA = 8 (one CTA per tile), 4 or 2: a tile is cut into 8 / WARPS horizontal strips,
// each composited by its own CTA with its own staging pipeline (batches of WARPS * 32 records) and its
// own early exit.  Why: per batch a CTA waits for its slowest warp, and the slowest warp is the one
// over the densest 4x2 block (CPU work model tools/analyze_balance.py: the other warps wait 51 % of the
// time, whatever the block -> warp assignment) -- so what keeps the schedulers fed is the NUMBER of
// independent CTAs resident on the SM, not the warps of one CTA.
__constant__ int sas_variant_flags_dev = 0xff;

// PACKED (WARPS = 4): the tile's groups that link Gaussians can reach (all groups for a view tile) are packed densely,
// four per warp, in sas_group_bit order, and the tile becomes ONE CTA of 4 warps (up to 16 packed groups) or TWO (more
// than 16): a tile with 14 reachable groups keeps 4 warps busy instead of leaving most of 8 warps waiting at the
// batch barrier while they hold warp slots and registers.  One CTA shape, 16 resident per SM.
//
// VREC: background records come from `vrec`, the view buckets' records materialised in depth order with exact reach masks
// (sas_bin.cu bg_reach_kernel): a merged entry with the top bit set is a position in vrec, a view tile's entry e is
// vrec[bucket offset + e] (no id list at all), and such a record's mask is taken from its b.z instead of being computed.
template <int WARPS, int MODE, bool COUNT, bool PACKED = false, bool BULK = false, bool VREC = false>
__global__ void __launch_bounds__(WARPS * 32, MODE == 3 ? 10 : 64 / WARPS)
composite_kernel(const SasRecord *__restrict__ rec, const int *__restrict__ merged_gid, const uint2 *__restrict__ pairs,
                 const int *__restrict__ img_tile_offset, const int *__restrict__ row_tile_offset, const int *__restrict__ view_of_image,
                 const unsigned int *__restrict__ tile_dyn_mask, const int *__restrict__ comp_len, const SasRecord *__restrict__ vrec,
                 const int *__restrict__ merged2, const int *__restrict__ comp_list,
                 SasWork *work, int n_bg, int n_dyn, int C, int I, int W, int H, int tile_w, int tile_h, float bg0, float bg1, float bg2,
                 SasOutputs img_out, SasViewBuf view_out, long long cap_src, long long cap_all, SasCounters *counters, int max_items) {
    constexpr int NT = WARPS * 32;     // threads = records per batch
    constexpr int PARTS = 8 / WARPS;   // CTAs per tile (PACKED: at most; a part beyond the tile's packed groups is skipped)
    __shared__ SasRecord stage[2][NT];
    __shared__ unsigned s_reach[MODE >= 2 ? 2 : 1][MODE >= 2 ? NT : 1];
    __shared__ SplitScratch s_split[MODE == 3 ? WARPS : 1];   // MODE 3: per-warp alpha rows / colours / hit queue
    __shared__ float2 s_org[MODE == 3 ? WARPS : 1][4];        // MODE 3: centre of the first pixel of each packed group
    __shared__ unsigned s_take;
    __shared__ __align__(8) unsigned long long s_mbar[2];     // BULK: one transaction barrier per stage buffer
    unsigned mbar_parity = 0u;                                // BULK: bit s = parity of the next completion of s_mbar[s]
    if (BULK) {
        if (threadIdx.x == 0) {
            sas_mbar_init(&s_mbar[0], 1u);
            sas_mbar_init(&s_mbar[1], 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    static_assert(MODE != 3 || (PACKED && WARPS == 4), "the split compositor is the packed 4-warp CTA");
    const int tid = threadIdx.x, lane = tid & 31;
    const int T = tile_w * tile_h;
    const long long IT = (long long)I * T;
    unsigned staged = 0;   // in eighths of a tile's list: a list staged by the two CTAs of a tile counts once
    // MODE 3: shared addresses of this lane's column of the alpha rows and of the colour rows of its warp's scratch
    const unsigned split_a_addr = MODE == 3 ? sas_pin((unsigned)__cvta_generic_to_shared(&s_split[MODE == 3 ? tid >> 5 : 0].alpha[0][lane & 7])) : 0u;
    const unsigned split_c_addr = MODE == 3 ? sas_pin((unsigned)__cvta_generic_to_shared(&s_split[MODE == 3 ? tid >> 5 : 0].col[0])) : 0u;

    // max_items: work items a CTA takes before it exits (0 = until the list is empty).  A bounded CTA hands its SM
    // slot back to the block scheduler, so kernels of another stream can move in while this one is still running.
    for (int taken = 0; max_items == 0 || taken < max_items; taken++) {
        __syncthreads();  // every thread is done with the previous tile's stage buffers and s_take
        if (tid == 0) s_take = atomicAdd(&work->comp_take, 1u);
        __syncthreads();
        const unsigned k = s_take;
        if (k >= work->comp_total * PARTS) break;
        const long long item = comp_list[k / PARTS];
        // Which 8x4 footprint of the tile this warp composites.  Hardware warp w of every resident CTA issues
        // from scheduler w % 4, and the background makes the same footprint of the same tile heavy in every
        // image of a view, so a fixed warp -> footprint map loads the four schedulers unevenly (profile: 0.90 M
        // to 1.35 M instructions per scheduler); rotating it with the work-item index evens that out.
        const int rot = (sas_variant_flags_dev & 128) ? (int)((k / PARTS) & 7u) : 0;
        const int warp = (((int)(k % PARTS) * WARPS + (tid >> 5)) + rot) & 7;

        // ---- resolve the item: id list (base, stride), record bases, output image ----
        const bool is_view = item >= IT;
        int out_img, tile, v;
        const int *gids;   // gid of entry e is gids[e * gstride]
        int gstride, n;
        unsigned vcode = 0u;   // VREC, view tile: entry e is vrec[vcode & 0x7fffffff + e]
        unsigned dyn_mask = 0xffffffffu;  // bit (warp*4 + group): a link Gaussian of this image can touch the group's 4x2 block
        if (is_view) {
            const long long vt = item - IT;
            v = (int)(vt / T); tile = (int)(vt - (long long)v * T);
            out_img = v;
            long long off = row_tile_offset[vt], end = row_tile_offset[vt + 1];
            if (off < 0 || end < off) { off = 0; end = 0; }
            if (end > cap_src) end = cap_src;
            if (off > end) off = end;
            gids = reinterpret_cast<const int *>(pairs + off) + 1;  // .y of each (depth, gid) pair
            gstride = 2;
            n = (int)(end - off);
            vcode = 0x80000000u | (unsigned)off;
        } else {
            out_img = (int)(item / T); tile = (int)(item - (long long)out_img * T);
            v = view_of_image[out_img];
            long long off = img_tile_offset[item], end = img_tile_offset[item + 1];
            if (off < 0 || end < off) { off = 0; end = 0; }
            if (end > cap_all) end = cap_all;
            if (off > end) off = end;
            // merged2 (VREC only): the tile's second CTA has a list of its own -- the entries that reach one of ITS groups
            const bool second = VREC && merged2 != nullptr && (k % PARTS) == 1u;
            gids = (second ? merged2 : merged_gid) + off;
            gstride = 1;
            n = (int)(end - off);
            if (comp_len) n = min(n, __ldg(comp_len + (second ? IT : 0ll) + item));  // the merge left out records no composited group can see
            dyn_mask = tile_dyn_mask[item];
        }
        // background records live per view, link records per image; a merged id below n_bg selects the former
        const SasRecord *rec_bg = rec + (size_t)v * n_bg;
        const SasRecord *rec_dy = rec + (size_t)C * n_bg + (size_t)(is_view ? 0 : out_img) * n_dyn - n_bg;
#define SAS_REC_OF(gid) (VREC ? ((gid) < 0 ? vrec + ((gid) & 0x7fffffff) : rec_dy + (gid)) : (((gid) < n_bg ? rec_bg : rec_dy) + (gid)))
#define SAS_GID_AT(e) ((VREC && is_view) ? (int)(vcode + (unsigned)(e)) : __ldg(gids + (size_t)(e) * gstride))

        // ---- this lane's pixel ----
        // warp = 8x4 footprint; its four 8-lane groups own the four 4x2 blocks of it (sas_common.cuh: sas_group_bit)
        const int ty = tile / tile_w, tx = tile - ty * tile_w;
        const int grp = lane >> 3, gl = lane & 7;
        int fx, fy, px, py, warp_arg = warp;
        bool group_static;
        int stage_weight = WARPS;
        if (PACKED) {
            const unsigned pm = dyn_mask & sas_inside_group_mask(tile, tile_w, W, H);
            const int ng = __popc(pm);
            // part p of the tile holds packed groups [p * WARPS * 4, (p + 1) * WARPS * 4); a part past the last group has
            // nothing to do (uniform over the CTA)
            const int part = (int)(k % PARTS);
            if (part > 0 && part * WARPS * 4 >= ng) continue;
            if (ng <= WARPS * 4) stage_weight = 8;                     // the only CTA of this tile
            const int kq = part * WARPS * 4 + (tid >> 5) * 4 + grp;   // this group's rank among the packed groups
            group_static = kq >= ng;
            const unsigned bit = group_static ? 0u : __fns(pm, 0u, kq + 1);  // its bit in the sas_group_bit layout
            const int cx = ((bit >> 2) & 1) * 2 + (bit & 1), ry = (bit >> 3) * 2 + ((bit >> 1) & 1);
            fx = tx * SAS_TILE_PX; fy = ty * SAS_TILE_PX;
            px = fx + 4 * cx + (gl & 3); py = fy + 2 * ry + (gl >> 2);
            warp_arg = (int)(__shfl_sync(0xffffffffu, bit, 0) | (__shfl_sync(0xffffffffu, bit, 8) << 8) |
                             (__shfl_sync(0xffffffffu, bit, 16) << 16) | (__shfl_sync(0xffffffffu, bit, 24) << 24));
            // published by the first batch barrier below; the barrier at the top of the item loop protects its reuse
            if (MODE == 3 && gl == 0) s_org[tid >> 5][grp] = make_float2((float)(fx + 4 * cx) + 0.5f, (float)(fy + 2 * ry) + 0.5f);
        } else {
            fx = tx * SAS_TILE_PX + (warp & 1) * 8; fy = ty * SAS_TILE_PX + (warp >> 1) * 4;  // 8 wide x 4 high
            px = fx + (grp & 1) * 4 + (gl & 3); py = fy + (grp >> 1) * 2 + (gl >> 2);
            // A group no link Gaussian of this image can touch is pixel-identical to the view's
            // background-only tile: it is skipped here and copied afterwards (copy_static_tiles_kernel).
            group_static = ((dyn_mask >> (warp * 4 + grp)) & 1u) == 0u;
        }
        const bool inside = (px < W) && (py < H);
        const float X0 = (float)fx + 0.5f, Y0 = (float)fy + 0.5f;
        Pixel p;
        p.pxf = (!inside || group_static) ? __int_as_float(0x7fc00000) : (float)px + 0.5f;
        p.pyf = (float)py + 0.5f;
        p.T = 1.0f; p.ar = p.ag = p.ab = p.ad = 0.0f;
        p.nc = 0;

        // ---- batches: prologue puts batch 0 in flight and the id of batch 1 in a register ----
        const int nb = (n + NT - 1) / NT;
        int gid_next = 0;
        if (BULK) {
            if (tid == 0 && n > 0) sas_mbar_expect_tx(&s_mbar[0], (unsigned)min(NT, n) * 48u);
            if (tid < n) { const int g0 = SAS_GID_AT(tid); sas_bulk_copy48(&stage[0][tid], SAS_REC_OF(g0), &s_mbar[0]); }
        } else {
            if (tid < n) { const int g0 = SAS_GID_AT(tid); stage_record(&stage[0][tid], SAS_REC_OF(g0)); }
            sas_cp_async_commit();
        }
        if (NT + tid < n) gid_next = SAS_GID_AT(NT + tid);
        for (int b = 0; b < nb; b++) {
            if (BULK) {   // every thread waits for the whole batch (all bytes landed: the barrier's transaction count is zero)
                sas_mbar_wait(&s_mbar[b & 1], (mbar_parity >> (b & 1)) & 1u);
                mbar_parity ^= 1u << (b & 1);
            } else {
                sas_cp_async_wait<0>();
            }
            if (MODE >= 2) {  // this thread's record of batch b has landed: which of the tile's 32 groups can it reach
                unsigned m = 0u;
                if (b * NT + tid < n) {
                    const float4 A = stage[b & 1][tid].a;
                    const float4 B = stage[b & 1][tid].b;
                    if (VREC && B.w < 0.0f) m = __float_as_uint(B.z);  // a view bucket record: exact mask, precomputed
                    else m = sas_group_reach_mask(sas_group_reach(make_float4(A.x - B.z, A.x + B.z, A.y - B.w, A.y + B.w)), tx, ty);
                }
                s_reach[b & 1][tid] = m;
            }
            if (__syncthreads_count(p.done()) == NT) break;  // also publishes batch b (records and reach masks) to the CTA
            const int cnt = min(NT, n - b * NT);
            staged += (unsigned)(cnt * stage_weight);
            const int buf = b & 1;
            {  // batch b+1 -> other buffer; id of batch b+2 -> register
                const int e1 = (b + 1) * NT + tid;
                if (BULK) {
                    const int c1 = min(NT, n - (b + 1) * NT);
                    if (tid == 0 && c1 > 0) sas_mbar_expect_tx(&s_mbar[buf ^ 1], (unsigned)c1 * 48u);
                    if (e1 < n) sas_bulk_copy48(&stage[buf ^ 1][tid], SAS_REC_OF(gid_next), &s_mbar[buf ^ 1]);
                } else {
                    if (e1 < n) stage_record(&stage[buf ^ 1][tid], SAS_REC_OF(gid_next));
                    sas_cp_async_commit();
                }
                const int e2 = e1 + NT;
                if (e2 < n) gid_next = SAS_GID_AT(e2);
            }
            if (__all_sync(0xffffffffu, p.done())) continue;  // whole warp finished: only keep feeding the pipeline
            if (MODE == 3) process_batch_split<COUNT>(stage[buf], s_reach[buf], (unsigned)warp_arg, cnt, p, s_split[MODE == 3 ? tid >> 5 : 0], s_org[MODE == 3 ? tid >> 5 : 0], split_a_addr, split_c_addr);
            else process_batch<MODE == 3 ? 2 : MODE, COUNT, PACKED>(stage[buf], s_reach[MODE >= 2 ? buf : 0], warp_arg, cnt, p, X0, Y0);
        }
        if (!BULK) sas_cp_async_wait<0>();
#undef SAS_REC_OF
#undef SAS_GID_AT

        if (inside && !group_static) {
            const size_t pi = ((size_t)out_img * H + py) * W + px;
            const float Aout = 1.0f - p.T;
            const float depth = p.ad / fmaxf(Aout, 1e-10f);
            if (is_view) {  // raw accumulators and T; the background is blended when the tile is copied into an image
                view_out.acc[4 * pi + 0] = p.ar; view_out.acc[4 * pi + 1] = p.ag; view_out.acc[4 * pi + 2] = p.ab;
                view_out.acc[4 * pi + 3] = p.T;
                view_out.depth[pi] = depth;
                view_out.n_contrib[pi] = p.nc;
            } else {
                write_pixel(img_out, pi, fmaf(p.T, bg0, p.ar), fmaf(p.T, bg1, p.ag), fmaf(p.T, bg2, p.ab), depth, Aout, p.nc);
            }
        }
    }
    if (tid == 0 && staged) atomicAdd(&counters->n_staged, (unsigned long long)staged);
}

// Images whose link bucket for a tile is empty take the tile from their view's buffers; the
// background blend and alpha = 1 - T are computed from the stored T exactly as the compositor does.
__global__ void __launch_bounds__(CB) copy_static_tiles_kernel(const unsigned int *__restrict__ row_tile_count, const int *__restrict__ view_of_image,
                                                               const unsigned int *__restrict__ tile_dyn_mask, int C, int W, int H, int tile_w,
                                                               int tile_h, float bg0, float bg1, float bg2, SasOutputs img_out,
                                                               SasViewBuf view_out) {
    const int tile = blockIdx.x, im = blockIdx.y;
    const int T = tile_w * tile_h;
    // tiles with link Gaussians were composited from their merged list, except the 4x2 groups none of them can touch
    const unsigned dyn = row_tile_count[(size_t)(C + im) * T + tile] != 0u ? tile_dyn_mask[(size_t)im * T + tile] : 0u;
    if (dyn == 0xffffffffu) return;
    const int v = view_of_image[im];
    const int ty = tile / tile_w, tx = tile - ty * tile_w;
    const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
    const int px = tx * SAS_TILE_PX + lx, py = ty * SAS_TILE_PX + ly;
    if (px >= W || py >= H) return;
    if ((dyn >> sas_group_bit(lx >> 2, ly >> 1)) & 1u) return;
    const size_t pi = ((size_t)im * H + py) * W + px, pv = ((size_t)v * H + py) * W + px;
    if (row_tile_count[(size_t)v * T + tile] == 0u) {  // nothing at all in this tile: T = 1, accumulators 0
        write_pixel(img_out, pi, fmaf(1.0f, bg0, 0.0f), fmaf(1.0f, bg1, 0.0f), fmaf(1.0f, bg2, 0.0f), 0.0f, 0.0f, 0);
        return;
    }
    const float4 a = reinterpret_cast<const float4 *>(view_out.acc)[pv];  // acc rgb, T
    write_pixel(img_out, pi, fmaf(a.w, bg0, a.x), fmaf(a.w, bg1, a.y), fmaf(a.w, bg2, a.z), view_out.depth[pv], 1.0f - a.w,
                view_out.n_contrib[pv]);
}

// splatfacto get_outputs post-processing (nerfstudio, [RECALLED]; reached from
// /root/reference/sim_a_splat/ns_utils/nerfstudio_utils.py:166-172)
__global__ void __launch_bounds__(256) ns_max_kernel(const float *__restrict__ depth, int P, unsigned int *__restrict__ img_max) {
    const int im = blockIdx.y;
    unsigned m = 0u;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < P; i += gridDim.x * 256) m = max(m, __float_as_uint(depth[(size_t)im * P + i]));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMax(img_max + im, m);  // depths are >= 0: uint order == float order
}

__global__ void __launch_bounds__(256) ns_post_kernel(float *__restrict__ rgb, float *__restrict__ depth, const float *__restrict__ alpha,
                                                      uint8_t *__restrict__ rgb_u8, uint16_t *__restrict__ depth_f16, int P, float bg0, float bg1,
                                                      float bg2, const unsigned int *__restrict__ img_max) {
    const int im = blockIdx.y;
    const float dmax = __uint_as_float(img_max[im]);
    for (int i = blockIdx.x * 256 + threadIdx.x; i < P; i += gridDim.x * 256) {
        const size_t pi = (size_t)im * P + i;
        const float a = alpha[pi];
        const float om = 1.0f - a;
        const float r = fminf(fmaxf(rgb[3 * pi + 0] + om * bg0, 0.0f), 1.0f);
        const float g = fminf(fmaxf(rgb[3 * pi + 1] + om * bg1, 0.0f), 1.0f);
        const float b = fminf(fmaxf(rgb[3 * pi + 2] + om * bg2, 0.0f), 1.0f);
        rgb[3 * pi + 0] = r; rgb[3 * pi + 1] = g; rgb[3 * pi + 2] = b;
        if (rgb_u8) { rgb_u8[3 * pi + 0] = to_u8(r); rgb_u8[3 * pi + 1] = to_u8(g); rgb_u8[3 * pi + 2] = to_u8(b); }
        if (depth && !(a > 0.0f)) {
            depth[pi] = dmax;
            if (depth_f16) depth_f16[pi] = __half_as_ushort(__float2half_rn(dmax));
        }
    }
}

int g_comp_sms = 148;
int g_comp_items = 1; // SAS_COMP_ITEMS: work items per compositor CTA before it exits (0 = persistent CTAs); measured on the
                      // pusht batch: 1 -> 1.208, 2 -> 1.213, 4 -> 1.221, 8 -> 1.235, persistent -> 1.231 ms/step
bool g_comp_items_env = false;
int g_comp_ctas = 8;  // persistent CTAs per SM of the default compositor (SAS_COMP_CTAS, experiments: 8 is best)

}  // namespace

cudaError_t sas_composite_init(int n_sms) {
    g_comp_sms = n_sms > 0 ? n_sms : 148;
    if (const char *e = getenv("SAS_COMP_ITEMS")) { g_comp_items = atoi(e) > 0 ? atoi(e) : 0; g_comp_items_env = true; }
    if (const char *e = getenv("SAS_COMP_CTAS")) g_comp_ctas = atoi(e) > 0 ? atoi(e) : 8;
    const int vf = sas_variant_flags();
    return cudaMemcpyToSymbol(sas_variant_flags_dev, &vf, sizeof(int));
}

// Registers are capped at 32 so that 64 warps stay resident per SM (measured 4.5 % faster than 48).
// SAS_VARIANT (sas_common.cuh): bit 0 per-group hit masks, bits 4-5 CTAs per tile.
cudaError_t sas_launch_composite(const SasRecord *rec, const int *merged_gid, const uint2 *pairs, const int *img_tile_offset,
                                 const int *row_tile_offset, const unsigned int *row_tile_count, const int *view_of_image,
                                 const unsigned int *tile_dyn_mask, const int *comp_len, const SasRecord *vrec, const int *merged2, const int *comp_list, SasWork *work, int n_bg,
                                 int n_dyn, int C, int I, int W, int H, int tile_w, int tile_h, float bg0, float bg1, float bg2,
                                 SasOutputs img_out, SasViewBuf view_out, long long cap_src, long long cap_all, SasCounters *counters,
                                 cudaStream_t st, int *n_launches) {
#define SAS_ARGS rec, merged_gid, pairs, img_tile_offset, row_tile_offset, view_of_image, tile_dyn_mask, comp_len, vrec, merged2, comp_list, work, n_bg, n_dyn, C, I, W, H, \
                 tile_w, tile_h, bg0, bg1, bg2, img_out, view_out, cap_src, cap_all, counters, max_items
    const int vf = sas_variant_flags();
    const int parts = (vf >> 4) & 3;  // 0: one CTA per tile, 1: two, 2 or 3: four
    // bounded CTAs (SAS_COMP_ITEMS > 0): enough of them for every possible item; otherwise persistent CTAs
    const long long all_items = (long long)(I + C) * tile_w * tile_h;
    // default: one item per CTA, more when the list could be far longer than the CTAs the GPU holds (most of those CTAs
    // would only find the list empty: 4 cameras x 1280x720 has 28 800 possible items and a few thousand real ones)
    const int max_items = g_comp_items_env ? g_comp_items : (int)max(1ll, (all_items + (long long)g_comp_sms * 64 - 1) / ((long long)g_comp_sms * 64));
#define SAS_GRID(W) (max_items > 0 ? (int)max((long long)g_comp_sms * (64 / (W)), (all_items * (8 / (W)) + max_items - 1) / max_items) \
                                   : g_comp_sms * (64 / (W)))
#define SAS_COMP(W, M) composite_kernel<W, M, true><<<SAS_GRID(W), W * 32, 0, st>>>(SAS_ARGS)
    const int mode = (vf & 1) ? ((vf & 64) ? 2 : 1) : 0;
    if (parts == 0) {
        if (mode == 2) {
            // the default: counts per-pixel contributors only when the caller asked for n_contrib
            const int grid = max_items > 0 ? SAS_GRID(8) : g_comp_sms * g_comp_ctas;
            if (vf & 0x200) {  // packed groups: every tile as one or two 4-warp CTAs of up to 16 packed groups each
                const int grid4 = max_items > 0 ? SAS_GRID(4) : g_comp_sms * 16;
                if (vf & 0x400) {  // split: alpha record-parallel, only the transmittance chain serial (MODE 3)
                    if (img_out.n_contrib) composite_kernel<4, 3, true, true><<<grid4, 128, 0, st>>>(SAS_ARGS);
                    else composite_kernel<4, 3, false, true><<<grid4, 128, 0, st>>>(SAS_ARGS);
                } else if (vf & 0x800) {  // bulk-async staging: cp.async.bulk + mbarrier expect_tx instead of LDGSTS + wait_group
                    if (img_out.n_contrib) composite_kernel<4, 2, true, true, true><<<grid4, 128, 0, st>>>(SAS_ARGS);
                    else composite_kernel<4, 2, false, true, true><<<grid4, 128, 0, st>>>(SAS_ARGS);
                } else if (vrec) {  // background records from the view buckets' contiguous copies, exact masks (bit 13)
                    if (img_out.n_contrib) composite_kernel<4, 2, true, true, false, true><<<grid4, 128, 0, st>>>(SAS_ARGS);
                    else composite_kernel<4, 2, false, true, false, true><<<grid4, 128, 0, st>>>(SAS_ARGS);
                } else if (img_out.n_contrib) composite_kernel<4, 2, true, true><<<grid4, 128, 0, st>>>(SAS_ARGS);
                else composite_kernel<4, 2, false, true><<<grid4, 128, 0, st>>>(SAS_ARGS);
            } else if (img_out.n_contrib) composite_kernel<8, 2, true><<<grid, 256, 0, st>>>(SAS_ARGS);
            else composite_kernel<8, 2, false><<<grid, 256, 0, st>>>(SAS_ARGS);
        } else if (mode == 1) SAS_COMP(8, 1); else SAS_COMP(8, 0);
    } else if (parts == 1) {
        if (mode == 2) SAS_COMP(4, 2); else SAS_COMP(4, 1);
    } else {
        if (mode == 2) SAS_COMP(2, 2); else SAS_COMP(2, 1);
    }
#undef SAS_GRID
#undef SAS_COMP
#undef SAS_ARGS
    dim3 grid(tile_w * tile_h, I);
    copy_static_tiles_kernel<<<grid, CB, 0, st>>>(row_tile_count, view_of_image, tile_dyn_mask, C, W, H, tile_w, tile_h, bg0, bg1, bg2, img_out,
                                                  view_out);
    if (n_launches) *n_launches += 2;
    return cudaGetLastError();
}

cudaError_t sas_launch_ns_post(float *rgb, float *depth, const float *alpha, uint8_t *rgb_u8, uint16_t *depth_f16, int I, int W, int H,
                               float bg0, float bg1, float bg2, unsigned int *img_max, cudaStream_t st, int *n_launches) {
    const int P = W * H;
    dim3 grid(min((P + 255) / 256, 64), I);
    cudaError_t e = cudaMemsetAsync(img_max, 0, sizeof(unsigned int) * I, st);
    if (e != cudaSuccess) return e;
    if (depth) {
        ns_max_kernel<<<grid, 256, 0, st>>>(depth, P, img_max);
        if (n_launches) *n_launches += 1;
    }
    ns_post_kernel<<<grid, 256, 0, st>>>(rgb, depth, alpha, rgb_u8, depth_f16, P, bg0, bg1, bg2, img_max);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}
