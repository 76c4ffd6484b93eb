// sas_composite.cu -- K5: per-tile front-to-back alpha compositing of RGB + depth + alpha
// (gsplat 1.5.2 rasterize_to_pixels, RGB+ED), and the splatfacto get_outputs post-processing.
//
// One CTA (256 threads) per 16x16 tile; each warp owns an 8x4 pixel footprint, one pixel per lane.
// The tile's depth-sorted Gaussian ids are consumed in batches of 256: every thread gathers one
// 48 B projected record into shared memory with three 16 B cp.async copies (double buffered, the
// id for the batch after next prefetched into a register), so global latency overlaps the
// arithmetic of the current batch.  Inside a batch each warp first tests 32 Gaussians at once
// (one per lane) against its footprint using the record's screen-space radii, ballots the hits,
// and only walks the set bits -- in order, so front-to-back order is preserved.  A lane stops
// when T*(1-alpha) <= 1e-4; a warp skips once all its lanes stopped; the CTA leaves the batch
// loop when __syncthreads_count says all 256 pixels stopped.
//
// The footprint test is conservative: a pixel outside a Gaussian's radius box has
// sigma > 0.5*3.33^2 = 5.544 > ln(255), i.e. alpha < 1/255 for any opacity <= 1, so the reference
// rule would have skipped it as well (one extra pixel of slack absorbs rounding).  Likewise
// sigma > SAS_SIGMA_SKIP skips the exp.  Neither changes any output bit.
//
// Not bandwidth bound: FP32 issue bound (sigma, the polynomial exp, 4 FMA accumulations).
#include "sas_common.cuh"

namespace {

constexpr int CB = 256;  // Gaussians per batch = threads per CTA

__device__ __forceinline__ unsigned char to_u8(float v) {
    v = fminf(fmaxf(v, 0.0f), 1.0f);
    return (unsigned char)(v * 255.0f + 0.5f);
}

__global__ void __launch_bounds__(CB) composite_kernel(const SasRecord *__restrict__ rec, const int *__restrict__ sorted_gid,
                                                       const int *__restrict__ tile_offset, const int *__restrict__ view_of_image,
                                                       int n_bg, int n_dyn, int C, int W, int H, int tile_w, int tile_h,
                                                       float bg0, float bg1, float bg2, float *__restrict__ out_rgb,
                                                       float *__restrict__ out_depth, float *__restrict__ out_alpha,
                                                       uint8_t *__restrict__ out_u8, int *__restrict__ out_nc, long long max_isects,
                                                       SasCounters *counters) {
    __shared__ SasRecord stage[2][CB];

    const int tile = blockIdx.x, im = blockIdx.y;
    const int ty = tile / tile_w, tx = tile - ty * tile_w;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t tix = (size_t)im * tile_w * tile_h + tile;
    long long off = tile_offset[tix], end = tile_offset[tix + 1];
    if (off < 0 || end < off) { off = 0; end = 0; }
    if (end > max_isects) end = max_isects;
    if (off > end) off = end;
    const int n = (int)(end - off);
    const int *gids = sorted_gid + off;
    // background records live per view, link records per image (sas_common.cuh); a merged id below
    // n_bg selects the former.  rec_of(gid) = the record of canonical Gaussian gid for this image.
    const SasRecord *rec_bg = rec + (size_t)view_of_image[im] * n_bg;
    const SasRecord *rec_dy = rec + (size_t)C * n_bg + (size_t)im * n_dyn - n_bg;
#define SAS_REC_OF(gid) (((gid) < n_bg ? rec_bg : rec_dy) + (gid))

    // warp footprint: 8 wide x 4 high
    const int fx = tx * SAS_TILE_PX + (warp & 1) * 8, fy = ty * SAS_TILE_PX + (warp >> 1) * 4;
    const int px = fx + (lane & 7), py = fy + (lane >> 3);
    const bool inside = (px < W) && (py < H);
    const float pxf = (float)px + 0.5f, pyf = (float)py + 0.5f;
    // pixel-centre bounds of the footprint, widened by one pixel of slack
    const float bx0 = (float)fx + 0.5f - 1.0f, bx1 = (float)fx + 7.5f + 1.0f;
    const float by0 = (float)fy + 0.5f - 1.0f, by1 = (float)fy + 3.5f + 1.0f;

    float T = 1.0f, ar = 0.0f, ag = 0.0f, ab = 0.0f, ad = 0.0f;
    int nc = 0;
    bool done = !inside;

    const int nb = (n + CB - 1) / CB;
    // prologue: batch 0 in flight, id of batch 1 in a register
    int gid_next = 0;
    if (tid < n) {
        const int g0 = __ldg(gids + tid);
        const SasRecord *s = SAS_REC_OF(g0);
        sas_cp_async16(&stage[0][tid].a, &s->a);
        sas_cp_async16(&stage[0][tid].b, &s->b);
        sas_cp_async16(&stage[0][tid].c, &s->c);
    }
    sas_cp_async_commit();
    if (CB + tid < n) gid_next = __ldg(gids + CB + tid);

    int staged = 0;
    for (int b = 0; b < nb; b++) {
        sas_cp_async_wait<0>();
        if (__syncthreads_count(done) == CB) break;  // also publishes batch b to the CTA
        staged += min(CB, n - b * CB);
        const int buf = b & 1;
        {  // batch b+1 -> other buffer; id of batch b+2 -> register
            const int e1 = (b + 1) * CB + tid;
            if (e1 < n) {
                const SasRecord *s = SAS_REC_OF(gid_next);
                sas_cp_async16(&stage[buf ^ 1][tid].a, &s->a);
                sas_cp_async16(&stage[buf ^ 1][tid].b, &s->b);
                sas_cp_async16(&stage[buf ^ 1][tid].c, &s->c);
            }
            sas_cp_async_commit();
            const int e2 = e1 + CB;
            if (e2 < n) gid_next = __ldg(gids + e2);
        }
        if (__all_sync(0xffffffffu, done)) continue;  // whole warp finished: only keep feeding the pipeline
        const int cnt = min(CB, n - b * CB);
        const SasRecord *sb = stage[buf];
        for (int j0 = 0; j0 < cnt; j0 += 32) {
            bool hit = false;
            if (j0 + lane < cnt) {
                const float4 A = sb[j0 + lane].a;
                const float4 B = sb[j0 + lane].b;
                hit = (A.x + B.z >= bx0) && (A.x - B.z <= bx1) && (A.y + B.w >= by0) && (A.y - B.w <= by1);
            }
            unsigned mask = __ballot_sync(0xffffffffu, hit);
            while (mask) {
                const int j = j0 + __ffs(mask) - 1;
                mask &= mask - 1;
                const float4 A = sb[j].a;
                const float4 B = sb[j].b;
                const float dx = A.x - pxf, dy = A.y - pyf;
                const float sigma = fmaf(fmaf(A.z, dx, A.w * dy), dx, (B.x * dy) * dy);
                if (!done && sigma >= 0.0f && sigma <= SAS_SIGMA_SKIP) {
                    const float al = fminf(SAS_ALPHA_MAX, B.y * sas_exp_neg(-sigma));
                    if (al >= SAS_ALPHA_MIN) {
                        const float nT = T * (1.0f - al);
                        if (nT <= SAS_T_MIN) {
                            done = true;
                        } else {
                            const float4 Cc = sb[j].c;
                            const float vis = al * T;
                            ar = fmaf(Cc.x, vis, ar);
                            ag = fmaf(Cc.y, vis, ag);
                            ab = fmaf(Cc.z, vis, ab);
                            ad = fmaf(Cc.w, vis, ad);
                            T = nT;
                            nc++;
                        }
                    }
                }
            }
            if (__all_sync(0xffffffffu, done)) break;
        }
    }
#undef SAS_REC_OF
    sas_cp_async_wait<0>();
    if (tid == 0 && staged) atomicAdd(&counters->n_staged, (unsigned long long)staged);

    if (inside) {
        const size_t pi = ((size_t)im * H + py) * W + px;
        const float Aout = 1.0f - T;
        const float r = fmaf(T, bg0, ar), g = fmaf(T, bg1, ag), bl = fmaf(T, bg2, ab);
        if (out_rgb) { out_rgb[3 * pi + 0] = r; out_rgb[3 * pi + 1] = g; out_rgb[3 * pi + 2] = bl; }
        if (out_alpha) out_alpha[pi] = Aout;
        if (out_depth) out_depth[pi] = ad / fmaxf(Aout, 1e-10f);
        if (out_u8) { out_u8[3 * pi + 0] = to_u8(r); out_u8[3 * pi + 1] = to_u8(g); out_u8[3 * pi + 2] = to_u8(bl); }
        if (out_nc) out_nc[pi] = nc;
    }
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
                                                      uint8_t *__restrict__ rgb_u8, int P, float bg0, float bg1, float bg2,
                                                      const unsigned int *__restrict__ img_max) {
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
        if (depth && !(a > 0.0f)) depth[pi] = dmax;
    }
}

}  // namespace

cudaError_t sas_launch_composite(const SasRecord *rec, const int *merged_gid, const int *img_tile_offset, const int *view_of_image,
                                 int n_bg, int n_dyn, int C, int I, int W, int H, int tile_w, int tile_h, float bg0, float bg1,
                                 float bg2, float *rgb, float *depth, float *alpha, uint8_t *rgb_u8, int *n_contrib,
                                 long long cap_all, SasCounters *counters, cudaStream_t st) {
    dim3 grid(tile_w * tile_h, I);
    composite_kernel<<<grid, CB, 0, st>>>(rec, merged_gid, img_tile_offset, view_of_image, n_bg, n_dyn, C, W, H, tile_w, tile_h, bg0,
                                          bg1, bg2, rgb, depth, alpha, rgb_u8, n_contrib, cap_all, counters);
    return cudaGetLastError();
}

cudaError_t sas_launch_ns_post(float *rgb, float *depth, const float *alpha, uint8_t *rgb_u8, int I, int W, int H, float bg0,
                               float bg1, float bg2, unsigned int *img_max, cudaStream_t st, int *n_launches) {
    const int P = W * H;
    dim3 grid(min((P + 255) / 256, 64), I);
    cudaError_t e = cudaMemsetAsync(img_max, 0, sizeof(unsigned int) * I, st);
    if (e != cudaSuccess) return e;
    if (depth) {
        ns_max_kernel<<<grid, 256, 0, st>>>(depth, P, img_max);
        if (n_launches) *n_launches += 1;
    }
    ns_post_kernel<<<grid, 256, 0, st>>>(rgb, depth, alpha, rgb_u8, P, bg0, bg1, bg2, img_max);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}
