// sas_project.cu -- K1: fused link-pose update + 3D covariance + EWA projection + SH colour
// + tile rectangle + per-(image,tile) counting, one thread per (image, Gaussian).
//
// Replaces (semantics): the per-link node transform viser applies after
// SplatHandler.draw_handler (/root/reference/sim_a_splat/splat/splat_handler.py:283-288) followed by
// gsplat 1.5.2 fully_fused_projection + spherical_harmonics + isect_tiles(count) reached from
// /root/reference/sim_a_splat/ns_utils/nerfstudio_utils.py:166-172.
//
// Layout: scene is structure-of-arrays in HBM -- float4 (mean.xyz, opacity), float4 quat wxyz,
// float4 (scale.xyz, link_id bits), and SH as (K*3) coefficient planes of n_pad floats, so every
// warp load is one fully used 128 B (planes) or 512 B (float4) request.  Bounded by HBM/L2
// bandwidth: 240 B read per Gaussian at SH degree 3, 8 B + 48 B (visible only) written.
//
// The arithmetic below is the same operation sequence as oracle/splat_oracle.c:project_one
// (written separately; see the numerics contract in sas_common.cuh).
#include "sas_common.cuh"

#include <cstdlib>

namespace {

__device__ __forceinline__ void sh_bases(int deg, float x, float y, float z, float *b) {
    b[0] = 0.2820947917738781f;
    if (deg < 1) return;
    float fTmpA = -0.48860251190292f;
    b[2] = -fTmpA * z; b[3] = fTmpA * x; b[1] = fTmpA * y;
    if (deg < 2) return;
    float z2 = z * z;
    float fTmpB = -1.092548430592079f * z;
    fTmpA = 0.5462742152960395f;
    float fC1 = x * x - y * y;
    float fS1 = 2.0f * x * y;
    b[6] = 0.9461746957575601f * z2 - 0.3153915652525201f;
    b[7] = fTmpB * x; b[5] = fTmpB * y; b[8] = fTmpA * fC1; b[4] = fTmpA * fS1;
    if (deg < 3) return;
    float fTmpC = -2.285228997322329f * z2 + 0.4570457994644658f;
    fTmpB = 1.445305721320277f * z;
    fTmpA = -0.5900435899266435f;
    float fC2 = x * fC1 - y * fS1;
    float fS2 = x * fS1 + y * fC1;
    b[12] = z * (1.865881662950577f * z2 - 1.119528997770346f);
    b[13] = fTmpC * x; b[11] = fTmpC * y; b[14] = fTmpB * fC1; b[10] = fTmpB * fS1;
    b[15] = fTmpA * fC2; b[9] = fTmpA * fS2;
}

// Projects Gaussian g for one row (camera [+ env]); writes its bin and, if visible, its record and
// radii.  Returns visibility and the tile rectangle.
template <int DEG>
__device__ __forceinline__ bool project_one(const float4 *__restrict__ mean_op, const float4 *__restrict__ quat,
                                            const float4 *__restrict__ scale_link, const float *__restrict__ sh_planes, int n_pad, int g,
                                            size_t ig, const float *__restrict__ link_poses, int L, const SasCamera *__restrict__ cam, int W,
                                            int H, int tile_w, int tile_h, SasRecord *__restrict__ rec, uint2 *__restrict__ bins,
                                            uint2 *__restrict__ radii_out, int &x0, int &y0, int &x1, int &y1, float4 &reach) {

    const float4 mo = __ldg(mean_op + g);
    const float4 qq = __ldg(quat + g);
    const float4 sl = __ldg(scale_link + g);
    const int link = __float_as_int(sl.w);

    // quaternion -> rotation (gsplat _quat_to_rotmat on the normalised quaternion)
    float Rg[9];
    {
        float w = qq.x, x = qq.y, y = qq.z, z = qq.w;
        float n = sqrtf(((w * w + x * x) + y * y) + z * z);
        n = fmaxf(n, 1e-12f);
        float inv = 1.0f / n;
        w = w * inv; x = x * inv; y = y * inv; z = z * inv;
        float x2 = x * x, y2 = y * y, z2 = z * z;
        float xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
        Rg[0] = 1.0f - 2.0f * (y2 + z2); Rg[1] = 2.0f * (xy - wz);        Rg[2] = 2.0f * (xz + wy);
        Rg[3] = 2.0f * (xy + wz);        Rg[4] = 1.0f - 2.0f * (x2 + z2); Rg[5] = 2.0f * (yz - wx);
        Rg[6] = 2.0f * (xz - wy);        Rg[7] = 2.0f * (yz + wx);        Rg[8] = 1.0f - 2.0f * (x2 + y2);
    }
    float mx = mo.x, my = mo.y, mz = mo.z;
    float R[9];
    float P[12];
    const bool moved = (link >= 0) && (link_poses != nullptr);
    if (moved) {  // S0: mu' = rot*mu + pos, R' = rot*R(q)
        const float *pp = link_poses + ((size_t)cam->env * L + link) * 12;
#pragma unroll
        for (int i = 0; i < 12; i++) P[i] = __ldg(pp + i);
        float tx = ((P[0] * mx + P[1] * my) + P[2] * mz) + P[9];
        float ty = ((P[3] * mx + P[4] * my) + P[5] * mz) + P[10];
        float tz = ((P[6] * mx + P[7] * my) + P[8] * mz) + P[11];
        mx = tx; my = ty; mz = tz;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++)
                R[3 * i + j] = (P[3 * i] * Rg[j] + P[3 * i + 1] * Rg[3 + j]) + P[3 * i + 2] * Rg[6 + j];
    } else {
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = Rg[i];
    }

    float vm[12];
#pragma unroll
    for (int i = 0; i < 12; i++) vm[i] = cam->vm[i];
    const float cxm = ((vm[0] * mx + vm[1] * my) + vm[2] * mz) + vm[3];
    const float cym = ((vm[4] * mx + vm[5] * my) + vm[6] * mz) + vm[7];
    const float czm = ((vm[8] * mx + vm[9] * my) + vm[10] * mz) + vm[11];

    bool vis = !(czm < SAS_NEAR || czm > SAS_FAR);
    float a = 0.f, b = 0.f, c = 0.f, det = 0.f, m2x = 0.f, m2y = 0.f, rx = 0.f, ry = 0.f;
    if (vis) {
        // S1: Sigma = (R S)(R S)^T
        float Ms[9], Sig[9];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            Ms[3 * i + 0] = R[3 * i + 0] * sl.x; Ms[3 * i + 1] = R[3 * i + 1] * sl.y; Ms[3 * i + 2] = R[3 * i + 2] * sl.z;
        }
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = i; j < 3; j++) {
                float v = (Ms[3 * i] * Ms[3 * j] + Ms[3 * i + 1] * Ms[3 * j + 1]) + Ms[3 * i + 2] * Ms[3 * j + 2];
                Sig[3 * i + j] = v; Sig[3 * j + i] = v;
            }
        // camera-frame covariance Rc Sigma Rc^T (Rc = rows of vm)
        float Tm[9], Cc[9];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++)
                Tm[3 * i + j] = (vm[4 * i] * Sig[j] + vm[4 * i + 1] * Sig[3 + j]) + vm[4 * i + 2] * Sig[6 + j];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++)
                Cc[3 * i + j] = (Tm[3 * i] * vm[4 * j] + Tm[3 * i + 1] * vm[4 * j + 1]) + Tm[3 * i + 2] * vm[4 * j + 2];

        // S2: EWA perspective projection (gsplat _persp_proj), blur, conic, radii, culls
        const float fx = cam->fx, fy = cam->fy, pcx = cam->cx, pcy = cam->cy;
        const float lim_x_pos = cam->lim_x_pos, lim_x_neg = cam->lim_x_neg;
        const float lim_y_pos = cam->lim_y_pos, lim_y_neg = cam->lim_y_neg;
        const float rz = 1.0f / czm;
        const float rz2 = rz * rz;
        const float txc = czm * fminf(lim_x_pos, fmaxf(-lim_x_neg, cxm * rz));
        const float tyc = czm * fminf(lim_y_pos, fmaxf(-lim_y_neg, cym * rz));
        const float ja = fx * rz, jb = -(fx * txc) * rz2, jc = fy * rz, jd = -(fy * tyc) * rz2;
        const float jc00 = ja * Cc[0] + jb * Cc[6], jc01 = ja * Cc[1] + jb * Cc[7], jc02 = ja * Cc[2] + jb * Cc[8];
        const float jc11 = jc * Cc[4] + jd * Cc[7], jc12 = jc * Cc[5] + jd * Cc[8];
        a = jc00 * ja + jc02 * jb;
        b = jc01 * jc + jc02 * jd;
        c = jc11 * jc + jc12 * jd;
        m2x = (fx * cxm) * rz + pcx;
        m2y = (fy * cym) * rz + pcy;
        a = a + SAS_EPS2D; c = c + SAS_EPS2D;
        det = a * c - b * b;
        vis = det > 0.0f;
        if (vis) {
            det = fmaxf(det, 1e-10f);
            rx = ceilf(SAS_EXTEND * sqrtf(a));
            ry = ceilf(SAS_EXTEND * sqrtf(c));
            if (rx <= 0.0f && ry <= 0.0f) vis = false;
            if (m2x + rx <= 0.0f || m2x - rx >= (float)W || m2y + ry <= 0.0f || m2y - ry >= (float)H) vis = false;
        }
    }
    if (!vis) {
        bins[ig] = make_uint2(0u, 0u);
        return false;
    }

    // S4: tile rectangle (gsplat _isect_tiles)
    const float tmx = m2x / (float)SAS_TILE_PX, tmy = m2y / (float)SAS_TILE_PX;
    const float trx = rx / (float)SAS_TILE_PX, try_ = ry / (float)SAS_TILE_PX;
    x0 = (int)floorf(tmx - trx); y0 = (int)floorf(tmy - try_);
    x1 = (int)ceilf(tmx + trx); y1 = (int)ceilf(tmy + try_);
    x0 = min(max(x0, 0), tile_w); x1 = min(max(x1, 0), tile_w);
    y0 = min(max(y0, 0), tile_h); y1 = min(max(y1, 0), tile_h);
    bins[ig] = make_uint2((unsigned)x0 | ((unsigned)y0 << 8) | ((unsigned)x1 << 16) | ((unsigned)y1 << 24), __float_as_uint(czm));
    radii_out[ig] = make_uint2((unsigned)fminf(rx, 2147483520.0f), (unsigned)fminf(ry, 2147483520.0f));  // saturating, as the oracle

    // S3: SH colour, view direction rotated into the link's rest frame
    float dx = mx - cam->campos[0], dy = my - cam->campos[1], dz = mz - cam->campos[2];
    if (moved) {
        float ux = (P[0] * dx + P[3] * dy) + P[6] * dz;
        float uy = (P[1] * dx + P[4] * dy) + P[7] * dz;
        float uz = (P[2] * dx + P[5] * dy) + P[8] * dz;
        dx = ux; dy = uy; dz = uz;
    }
    float dn = sqrtf((dx * dx + dy * dy) + dz * dz);
    dn = fmaxf(dn, 1e-12f);
    const float dinv = 1.0f / dn;
    dx = dx * dinv; dy = dy * dinv; dz = dz * dinv;
    constexpr int KN = (DEG + 1) * (DEG + 1);
    float bas[KN];
    sh_bases(DEG, dx, dy, dz, bas);
    float col[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        float acc = bas[0] * __ldg(sh_planes + (size_t)ch * n_pad + g);
#pragma unroll
        for (int k = 1; k < KN; k++) acc = acc + bas[k] * __ldg(sh_planes + (size_t)(3 * k + ch) * n_pad + g);
        col[ch] = fmaxf(acc + 0.5f, 0.0f);
    }

    // Screen-space extent the compositor culls with: a pixel can pass alpha >= 1/255 only if
    // sigma <= tau = ln(255*opacity), and sigma >= dx^2/(2a) (a = blurred cov_xx), so |dx| <= sqrt(2*tau*a).
    // Not part of the numerics contract (it only skips work): logf is the device libm, and the bound
    // carries a margin (tau + 0.01, x1.001, + 0.02 px) far above the rounding of sigma; it never exceeds
    // the gsplat radius + 1.
    const float tau = fmaxf(logf(255.0f * mo.w) + 0.01f, 0.0f);
    const float ex = fminf(sqrtf(2.0f * tau * a) * 1.001f + 0.02f, rx + 1.0f);
    const float ey = fminf(sqrtf(2.0f * tau * c) * 1.001f + 0.02f, ry + 1.0f);

    SasRecord r;
    r.a = make_float4(m2x, m2y, 0.5f * (c / det), -b / det);
    r.b = make_float4(0.5f * (a / det), mo.w, ex, ey);
    r.c = make_float4(col[0], col[1], col[2], czm);
    rec[ig] = r;
    reach = make_float4(m2x - ex, m2x + ex, m2y - ey, m2y + ey);
    return true;
}


// One thread per (row, Gaussian) of a span: the background span runs once per VIEW (rows = distinct
// cameras), the link span once per IMAGE (rows = env x camera) -- see sas_common.cuh.  Per-(row,tile)
// counts are accumulated in shared memory per CTA (one global RED per touched tile per CTA instead
// of one per intersection); images with more tiles than fit in shared memory add to the global
// counters directly.
template <int DEG, int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS)
project_kernel(const float4 *__restrict__ mean_op, const float4 *__restrict__ quat, const float4 *__restrict__ scale_link,
               const float *__restrict__ sh_planes, int n_pad, SasRowSpan span, const float *__restrict__ link_poses, int L,
               const SasCamera *__restrict__ row_cams, int W, int H, int tile_w, int tile_h, SasRecord *__restrict__ rec,
               uint2 *__restrict__ bins, uint2 *__restrict__ radii_out, unsigned int *__restrict__ row_tile_count,
               unsigned int *__restrict__ tile_dyn_mask, SasCounters *counters, int smem_tiles, int smem_masks) {
    extern __shared__ unsigned int s_cnt[];  // [smem_tiles] counts (0 = direct global counting), then [smem_tiles] group masks if smem_masks
    unsigned int *s_mask = s_cnt + smem_tiles;
    const int T = tile_w * tile_h;
    for (int t = threadIdx.x; t < smem_tiles * (1 + smem_masks); t += 256) s_cnt[t] = 0u;
    if (smem_tiles) __syncthreads();

    const int local = blockIdx.x * 256 + threadIdx.x;
    const int row = blockIdx.y;
    bool vis = false;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    float4 reach = make_float4(0.f, 0.f, 0.f, 0.f);
    if (local < span.n)
        vis = project_one<DEG>(mean_op, quat, scale_link, sh_planes, n_pad, span.g0 + local,
                               (size_t)span.rec0 + (size_t)row * span.n + local, link_poses, L, row_cams + span.row0 + row, W, H, tile_w,
                               tile_h, rec, bins, radii_out, x0, y0, x1, y1, reach);
    unsigned int *tc = row_tile_count + (size_t)(span.row0 + row) * T;
    // link span only (rows = images): per-(image,tile) mask of the 4x2 groups a link Gaussian can reach
    unsigned int *dm = tile_dyn_mask ? tile_dyn_mask + (size_t)row * T : nullptr;
    // If a Gaussian of the CTA covers many tiles, the CTA's (Gaussian, tile) intersections are counted as one
    // flattened list (block scan of the tile counts + a binary search per entry, as in emit_kernel), so a
    // screen-filling Gaussian does not leave 255 threads waiting for the one that owns it; otherwise every
    // thread counts its own few tiles.
    __shared__ unsigned int s_pre[257];
    __shared__ unsigned int s_rect[256];
    __shared__ unsigned int s_wsum[8];
    __shared__ SasGroupReach s_gr[256];
    const unsigned ntile = vis ? (unsigned)((x1 - x0) * (y1 - y0)) : 0u;
    const SasGroupReach my_gr = sas_group_reach(reach);
    auto count_one = [&](const SasGroupReach &gr, int x, int y) {
        const int t = y * tile_w + x;
        if (smem_tiles) atomicAdd(s_cnt + t, 1u); else atomicAdd(tc + t, 1u);
        if (dm) {
            const unsigned m = sas_group_reach_mask(gr, x, y);
            // read first: after the first few Gaussians of a CTA the bits are already set
            if (smem_masks) {
                if ((s_mask[t] & m) != m) atomicOr(s_mask + t, m);
            } else if ((__ldcg(dm + t) & m) != m) {
                atomicOr(dm + t, m);
            }
        }
    };
    if (__syncthreads_or(ntile > 16u)) {
        s_rect[threadIdx.x] = vis ? ((unsigned)x0 | ((unsigned)y0 << 8) | ((unsigned)x1 << 16) | ((unsigned)y1 << 24)) : 0u;
        if (dm) s_gr[threadIdx.x] = my_gr;
        unsigned total;
        {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            unsigned inc = ntile;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            if (lane == 31) s_wsum[warp] = inc;
            __syncthreads();
            unsigned pre = 0;
            total = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const unsigned sw = s_wsum[w];
                if (w < warp) pre += sw;
                total += sw;
            }
            s_pre[threadIdx.x] = pre + inc - ntile;
            if (threadIdx.x == 0) s_pre[256] = total;
            __syncthreads();
        }
        for (unsigned k = threadIdx.x; k < total; k += 256) {
            int lo = 0, hi = 256;  // largest g with s_pre[g] <= k
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_pre[mid] <= k) lo = mid; else hi = mid;
            }
            const unsigned bx = s_rect[lo], r = k - s_pre[lo];
            const int gx0 = bx & 0xff, gy0 = (bx >> 8) & 0xff, rw = (int)((bx >> 16) & 0xff) - gx0;
            const int dy = (int)r / rw;
            count_one(s_gr[lo], gx0 + ((int)r - dy * rw), gy0 + dy);
        }
    } else if (vis) {
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) count_one(my_gr, x, y);
    }
    {  // one counter update per warp instead of one per thread
        const unsigned m = __ballot_sync(0xffffffffu, vis);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counters->n_visible_src, (unsigned long long)__popc(m));
    }
    if (smem_tiles) {
        __syncthreads();
        for (int t = threadIdx.x; t < smem_tiles; t += 256) {
            const unsigned c = s_cnt[t];
            if (c) atomicAdd(tc + t, c);
            if (dm && smem_masks) {
                const unsigned m = s_mask[t];
                if (m) atomicOr(dm + t, m);
            }
        }
    }
}

}  // namespace

cudaError_t sas_launch_project(const float4 *mean_op, const float4 *quat, const float4 *scale_link, const float *sh_planes,
                               int n_pad, int sh_degree, const float *link_poses, int L, const SasCamera *row_cams,
                               SasRowSpan span, int W, int H, int tile_w, int tile_h, SasRecord *rec, uint2 *bins,
                               uint2 *radii_out, unsigned int *row_tile_count, unsigned int *tile_dyn_mask, SasCounters *counters,
                               cudaStream_t st) {
    if (span.n <= 0 || span.rows <= 0) return cudaSuccess;
    dim3 grid((span.n + 255) / 256, span.rows);
    const int T = tile_w * tile_h;
    const int smem_tiles = T <= SAS_SMEM_TILES_MAX ? T : 0;
    const int smem_masks = (tile_dyn_mask && smem_tiles && T <= SAS_SMEM_TILES_MAX / 2) ? 1 : 0;  // stays under the 48 KB default limit
    const size_t smem = (size_t)smem_tiles * (1 + smem_masks) * sizeof(unsigned int);
    // SAS_PROJECT=n (environment, read once): minimum resident CTAs/SM the register allocation targets (3, 4 or 5).
    static int variant = -1;
    if (variant < 0) {
        const char *e = getenv("SAS_PROJECT");
        variant = e ? atoi(e) : 4;
        if (variant < 3 || variant > 5) variant = 4;
    }
#define SAS_PROJ_ARGS mean_op, quat, scale_link, sh_planes, n_pad, span, link_poses, L, row_cams, W, H, tile_w, tile_h, rec, bins, radii_out, \
                      row_tile_count, tile_dyn_mask, counters, smem_tiles, smem_masks
#define SAS_PROJ(D)                                                                              \
    do {                                                                                         \
        if (variant == 3) project_kernel<D, 3><<<grid, 256, smem, st>>>(SAS_PROJ_ARGS);          \
        else if (variant == 5) project_kernel<D, 5><<<grid, 256, smem, st>>>(SAS_PROJ_ARGS);     \
        else project_kernel<D, 4><<<grid, 256, smem, st>>>(SAS_PROJ_ARGS);                       \
    } while (0)
    switch (sh_degree) {
        case 0: SAS_PROJ(0); break;
        case 1: SAS_PROJ(1); break;
        case 2: SAS_PROJ(2); break;
        default: SAS_PROJ(3); break;
    }
#undef SAS_PROJ
#undef SAS_PROJ_ARGS
    return cudaGetLastError();
}
