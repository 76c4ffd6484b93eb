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

// a0 b0 + a1 b1 + a2 b2 as one product and two fused multiply-adds, in this order (numerics contract: the oracle's
// dot3 / aff3 are the same chains); aff3 adds the products onto t.
__device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
    return fmaf(a2, b2, fmaf(a1, b1, a0 * b0));
}
__device__ __forceinline__ float aff3(float a0, float a1, float a2, float b0, float b1, float b2, float t) {
    return fmaf(a2, b2, fmaf(a1, b1, fmaf(a0, b0, t)));
}

// What a thread keeps in registers for ITS Gaussian while it walks a chunk of rows (envs x cameras of the link span,
// views of the background span): everything that does not depend on the row -- including the (DEG+1)^2 x 3 SH
// coefficients, so that they are read once per chunk of rows instead of once per (row, Gaussian) (SURVEY section 7:
// "loop cameras inside the thread that holds a Gaussian's 48 coefficients").
template <int DEG>
struct GaussConst {
    static constexpr int KN = (DEG + 1) * (DEG + 1);
    float mx, my, mz, op;
    float sx, sy, sz;
    int link;
    float Rg[9];
    float tau;          // ln(255 opacity) + margin (culling extents only)
    float ext;          // radius extend: 3.33, or the per-Gaussian opacity-aware factor (radius_rule = cuda_1_5)
};

// sh_s: this thread's column of the CTA's shared-memory coefficient block, coefficient i at sh_s[i * 256]
template <int DEG, bool SH_SMEM>
__device__ __forceinline__ void load_gauss(GaussConst<DEG> &G, const float4 *__restrict__ mean_op, const float4 *__restrict__ quat,
                                           const float4 *__restrict__ scale_link, const float *__restrict__ sh_planes, int n_pad, int g,
                                           float *__restrict__ sh_s, const float *__restrict__ extend) {
    const float4 mo = __ldg(mean_op + g);
    const float4 qq = __ldg(quat + g);
    const float4 sl = __ldg(scale_link + g);
    G.mx = mo.x; G.my = mo.y; G.mz = mo.z; G.op = mo.w;
    G.sx = sl.x; G.sy = sl.y; G.sz = sl.z;
    G.link = __float_as_int(sl.w);
    // quaternion -> rotation (gsplat _quat_to_rotmat on the normalised quaternion)
    float w = qq.x, x = qq.y, y = qq.z, z = qq.w;
    float n = sqrtf(fmaf(z, z, fmaf(y, y, fmaf(x, x, w * w))));
    n = fmaxf(n, 1e-12f);
    const float inv = 1.0f / n;
    w = w * inv; x = x * inv; y = y * inv; z = z * inv;
    const float x2 = x * x, y2 = y * y, z2 = z * z;
    const float xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
    G.Rg[0] = 1.0f - 2.0f * (y2 + z2); G.Rg[1] = 2.0f * (xy - wz);        G.Rg[2] = 2.0f * (xz + wy);
    G.Rg[3] = 2.0f * (xy + wz);        G.Rg[4] = 1.0f - 2.0f * (x2 + z2); G.Rg[5] = 2.0f * (yz - wx);
    G.Rg[6] = 2.0f * (xz - wy);        G.Rg[7] = 2.0f * (yz + wx);        G.Rg[8] = 1.0f - 2.0f * (x2 + y2);
    // Screen-space extent the compositor culls with (see project_row): not part of the numerics contract (it only
    // skips work): logf is the device libm, and the bound carries a margin far above the rounding of sigma.
    G.tau = fmaxf(logf(255.0f * mo.w) + 0.01f, 0.0f);
    G.ext = extend ? __ldg(extend + g) : SAS_EXTEND;
    if (SH_SMEM) {
#pragma unroll
        for (int i = 0; i < 3 * GaussConst<DEG>::KN; i++) sh_s[i * 256] = __ldg(sh_planes + (size_t)i * n_pad + g);
    }
}

// Projects the thread's Gaussian for one row (camera [+ env]); writes its bin and, if visible, its record and
// radii.  Returns visibility and the tile rectangle.  cam / pose: this row's camera block and (link Gaussians) the
// [rot | pos] block of the row's env and this Gaussian's link (both staged in shared memory by the CTA).
// SH_SMEM: the coefficients come from the thread's shared-memory column sh_s[i * 256] (staged once per chunk of rows);
// otherwise straight from the global coefficient planes sh_g[i * n_pad] (one row per CTA: nothing to reuse).
template <int DEG, bool SH_SMEM>
__device__ __forceinline__ bool project_row(const GaussConst<DEG> &G, const float *__restrict__ sh_s, const float *__restrict__ sh_g, int n_pad,
                                            size_t ig,
                                            const float *__restrict__ P /*12 floats or nullptr*/,
                                            const SasCamera *__restrict__ cam, int W, int H, int tile_w, int tile_h,
                                            SasRecord *__restrict__ rec, uint2 *__restrict__ bins, uint2 *__restrict__ radii_out, int &x0,
                                            int &y0, int &x1, int &y1, float4 &reach) {
    float mx = G.mx, my = G.my, mz = G.mz;
    float R[9];
    const bool moved = P != nullptr;
    if (moved) {  // S0: mu' = rot*mu + pos, R' = rot*R(q)
        const float tx = aff3(P[0], P[1], P[2], mx, my, mz, P[9]);
        const float ty = aff3(P[3], P[4], P[5], mx, my, mz, P[10]);
        const float tz = aff3(P[6], P[7], P[8], mx, my, mz, P[11]);
        mx = tx; my = ty; mz = tz;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) R[3 * i + j] = dot3(P[3 * i], P[3 * i + 1], P[3 * i + 2], G.Rg[j], G.Rg[3 + j], G.Rg[6 + j]);
    } else {
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = G.Rg[i];
    }

    float vm[12];
#pragma unroll
    for (int i = 0; i < 12; i++) vm[i] = cam->vm[i];
    const float cxm = aff3(vm[0], vm[1], vm[2], mx, my, mz, vm[3]);
    const float cym = aff3(vm[4], vm[5], vm[6], mx, my, mz, vm[7]);
    const float czm = aff3(vm[8], vm[9], vm[10], mx, my, mz, vm[11]);

    bool vis = !(czm < SAS_NEAR || czm > SAS_FAR);
    float a = 0.f, b = 0.f, c = 0.f, det = 0.f, m2x = 0.f, m2y = 0.f, rx = 0.f, ry = 0.f;
    if (vis) {
        // S1: Sigma = (R S)(R S)^T
        float Ms[9], Sig[9];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            Ms[3 * i + 0] = R[3 * i + 0] * G.sx; Ms[3 * i + 1] = R[3 * i + 1] * G.sy; Ms[3 * i + 2] = R[3 * i + 2] * G.sz;
        }
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = i; j < 3; j++) {
                const float v = dot3(Ms[3 * i], Ms[3 * i + 1], Ms[3 * i + 2], Ms[3 * j], Ms[3 * j + 1], Ms[3 * j + 2]);
                Sig[3 * i + j] = v; Sig[3 * j + i] = v;
            }
        // camera-frame covariance Rc Sigma Rc^T (Rc = rows of vm); entry [1][0] is never used
        float Tm[9], Cc[9];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) Tm[3 * i + j] = dot3(vm[4 * i], vm[4 * i + 1], vm[4 * i + 2], Sig[j], Sig[3 + j], Sig[6 + j]);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++)
                if (3 * i + j != 3) Cc[3 * i + j] = dot3(Tm[3 * i], Tm[3 * i + 1], Tm[3 * i + 2], vm[4 * j], vm[4 * j + 1], vm[4 * j + 2]);

        // S2: EWA perspective projection (gsplat _persp_proj), blur, conic, radii, culls
        const float fx = cam->fx, fy = cam->fy, pcx = cam->cx, pcy = cam->cy;
        const float rz = 1.0f / czm;
        const float rz2 = rz * rz;
        const float txc = czm * fminf(cam->lim_x_pos, fmaxf(-cam->lim_x_neg, cxm * rz));
        const float tyc = czm * fminf(cam->lim_y_pos, fmaxf(-cam->lim_y_neg, cym * rz));
        const float ja = fx * rz, jb = -(fx * txc) * rz2, jc = fy * rz, jd = -(fy * tyc) * rz2;
        const float jc00 = fmaf(jb, Cc[6], ja * Cc[0]), jc01 = fmaf(jb, Cc[7], ja * Cc[1]), jc02 = fmaf(jb, Cc[8], ja * Cc[2]);
        const float jc11 = fmaf(jd, Cc[7], jc * Cc[4]), jc12 = fmaf(jd, Cc[8], jc * Cc[5]);
        a = fmaf(jc02, jb, jc00 * ja);
        b = fmaf(jc02, jd, jc01 * jc);
        c = fmaf(jc12, jd, jc11 * jc);
        m2x = fmaf(fx * cxm, rz, pcx);
        m2y = fmaf(fy * cym, rz, pcy);
        a = a + SAS_EPS2D; c = c + SAS_EPS2D;
        det = fmaf(a, c, -(b * b));
        vis = det > 0.0f;
        if (vis) {
            det = fmaxf(det, 1e-10f);
            rx = ceilf(G.ext * sqrtf(a));
            ry = ceilf(G.ext * sqrtf(c));
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
        const float ux = dot3(P[0], P[3], P[6], dx, dy, dz);
        const float uy = dot3(P[1], P[4], P[7], dx, dy, dz);
        const float uz = dot3(P[2], P[5], P[8], dx, dy, dz);
        dx = ux; dy = uy; dz = uz;
    }
    float dn = sqrtf(dot3(dx, dy, dz, dx, dy, dz));
    dn = fmaxf(dn, 1e-12f);
    const float dinv = 1.0f / dn;
    dx = dx * dinv; dy = dy * dinv; dz = dz * dinv;
    constexpr int KN = (DEG + 1) * (DEG + 1);
    float bas[KN];
    sh_bases(DEG, dx, dy, dz, bas);
    float col[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        float acc = bas[0] * (SH_SMEM ? sh_s[ch * 256] : __ldg(sh_g + (size_t)ch * n_pad));
#pragma unroll
        for (int k = 1; k < KN; k++) acc = fmaf(bas[k], SH_SMEM ? sh_s[(3 * k + ch) * 256] : __ldg(sh_g + (size_t)(3 * k + ch) * n_pad), acc);
        col[ch] = fmaxf(acc + 0.5f, 0.0f);
    }

    // Screen-space extent the compositor culls with: a pixel can pass alpha >= 1/255 only if
    // sigma <= tau = ln(255*opacity), and sigma >= dx^2/(2a) (a = blurred cov_xx), so |dx| <= sqrt(2*tau*a).
    // The bound carries a margin (tau + 0.01, x1.001, + 0.02 px) far above the rounding of sigma; it never exceeds
    // the gsplat radius + 1.
    const float ex = fminf(sqrtf(2.0f * G.tau * a) * 1.001f + 0.02f, rx + 1.0f);
    const float ey = fminf(sqrtf(2.0f * G.tau * c) * 1.001f + 0.02f, ry + 1.0f);

    SasRecord r;
    r.a = make_float4(m2x, m2y, 0.5f * (c / det), -b / det);
    r.b = make_float4(0.5f * (a / det), G.op, ex, ey);
    r.c = make_float4(col[0], col[1], col[2], czm);
    rec[ig] = r;
    reach = make_float4(m2x - ex, m2x + ex, m2y - ey, m2y + ey);
    return true;
}

// One thread per Gaussian of a span, walking a chunk of ROWS: the background span runs once per VIEW (rows =
// distinct cameras), the link span once per IMAGE (rows = env x camera) -- see sas_common.cuh.  grid = (Gaussian
// blocks, row chunks).  The thread keeps its Gaussian's constants in registers and its SH coefficients in its own
// column of a shared-memory block across the chunk (read from HBM / L2 once per chunk instead of once per row); the
// CTA stages the chunk's cameras and link poses in shared memory.  Per-(row,tile) counts and group masks of the
// WHOLE chunk are accumulated in shared memory (one global RED per touched tile per CTA and row instead of one per
// intersection), so there is no barrier inside the row loop and the warps of a CTA drift through the rows
// independently; images with more tiles than fit in shared memory add to the global counters directly.
// (Round 2 first kept the 48 coefficients in registers with a barrier-fenced histogram per row: 120 registers, 2 CTAs
// per SM, ~30 % issue utilisation -- 0.140 ms against the 0.124 ms of the one-row-per-thread kernel it replaced.)
constexpr int PROJ_ROWS = 8;        // rows per chunk (fewer when the chunk's histograms would not fit shared memory)
constexpr int PROJ_MAX_LINKS = 8;   // link poses staged in shared memory up to this many links (else read from global)

template <int DEG, int MIN_CTAS, bool SH_SMEM>
__global__ void __launch_bounds__(256, MIN_CTAS)
project_kernel(const float4 *__restrict__ mean_op, const float4 *__restrict__ quat, const float4 *__restrict__ scale_link,
               const float *__restrict__ sh_planes, int n_pad, SasRowSpan span, const float *__restrict__ link_poses, int L,
               const SasCamera *__restrict__ row_cams, int W, int H, int tile_w, int tile_h, SasRecord *__restrict__ rec,
               uint2 *__restrict__ bins, uint2 *__restrict__ radii_out, unsigned int *__restrict__ row_tile_count,
               unsigned int *__restrict__ tile_dyn_mask, SasCounters *counters, int smem_tiles, int smem_masks, int rows_per_chunk,
               const float *__restrict__ extend) {
    // dynamic shared memory: (SH_SMEM) [3 KN][256] SH coefficients, then per row of the chunk [smem_tiles] counts (+ [smem_tiles] masks)
    extern __shared__ __align__(16) unsigned int s_dyn[];
    constexpr int KN3 = 3 * (DEG + 1) * (DEG + 1);
    float *sh_s = reinterpret_cast<float *>(s_dyn) + threadIdx.x;
    unsigned int *s_hist = s_dyn + (SH_SMEM ? KN3 * 256 : 0);
    const int hist_stride = smem_tiles * (1 + smem_masks);
    __shared__ SasCamera s_cam[PROJ_ROWS];
    __shared__ float s_pose[PROJ_ROWS][PROJ_MAX_LINKS][12];
    const int T = tile_w * tile_h;
    const int row_lo = blockIdx.y * rows_per_chunk, n_rows = min(rows_per_chunk, span.rows - row_lo);
    const bool poses_in_smem = link_poses != nullptr && L <= PROJ_MAX_LINKS;
    const int lane = threadIdx.x & 31;
    // stage the chunk's cameras (24 words each) and link poses; clear the chunk's histograms
    for (int k = threadIdx.x; k < n_rows * (int)(sizeof(SasCamera) / 4); k += 256) {
        const int rr = k / (int)(sizeof(SasCamera) / 4), w = k - rr * (int)(sizeof(SasCamera) / 4);
        reinterpret_cast<unsigned *>(&s_cam[rr])[w] = __ldg(reinterpret_cast<const unsigned *>(row_cams + span.row0 + row_lo + rr) + w);
    }
    for (int t = threadIdx.x; t < n_rows * hist_stride; t += 256) s_hist[t] = 0u;
    __syncthreads();
    if (poses_in_smem)
        for (int k = threadIdx.x; k < n_rows * L * 12; k += 256) {
            const int rr = k / (L * 12), w = k - rr * (L * 12);
            s_pose[rr][w / 12][w % 12] = __ldg(link_poses + (size_t)s_cam[rr].env * L * 12 + w);
        }

    const int local = blockIdx.x * 256 + threadIdx.x;
    const bool have = local < span.n;
    GaussConst<DEG> G;
    if (have) load_gauss<DEG, SH_SMEM>(G, mean_op, quat, scale_link, sh_planes, n_pad, span.g0 + local, sh_s, extend);
    else G.link = -1;
    const bool moved = have && G.link >= 0 && link_poses != nullptr;
    unsigned n_vis_thread = 0;
    __syncthreads();  // link poses are staged

    for (int rr = 0; rr < n_rows; rr++) {
        const int row = row_lo + rr;
        bool vis = false;
        int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
        float4 reach = make_float4(0.f, 0.f, 0.f, 0.f);
        if (have) {
            const float *P = !moved ? nullptr : poses_in_smem ? &s_pose[rr][G.link][0] : link_poses + ((size_t)s_cam[rr].env * L + G.link) * 12;
            vis = project_row<DEG, SH_SMEM>(G, sh_s, sh_planes + span.g0 + local, n_pad, (size_t)span.rec0 + (size_t)row * span.n + local, P,
                                            &s_cam[rr], W, H, tile_w, tile_h, rec, bins, radii_out, x0, y0, x1, y1, reach);
        }
        n_vis_thread += vis ? 1u : 0u;
        unsigned int *s_cnt = s_hist + rr * hist_stride, *s_mask = s_cnt + smem_tiles;
        unsigned int *tc = row_tile_count + (size_t)(span.row0 + row) * T;
        // link span only (rows = images): per-(image,tile) mask of the 4x2 groups a link Gaussian can reach
        unsigned int *dm = tile_dyn_mask ? tile_dyn_mask + (size_t)row * T : nullptr;
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
        // If a Gaussian of the WARP covers many tiles, the warp's (Gaussian, tile) intersections are counted as one
        // flattened list (warp scan of the tile counts + a binary search per entry through shuffles), so a
        // screen-filling Gaussian does not leave 31 lanes waiting for the one that owns it; otherwise every lane
        // counts its own few tiles.  Warp-level on purpose: no CTA barrier inside the row loop.
        if (__any_sync(0xffffffffu, ntile > 16u)) {
            unsigned inc = ntile;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            const unsigned total = __shfl_sync(0xffffffffu, inc, 31);
            const unsigned pre = inc - ntile;  // exclusive prefix of this lane
            const unsigned rect = vis ? ((unsigned)x0 | ((unsigned)y0 << 8) | ((unsigned)x1 << 16) | ((unsigned)y1 << 24)) : 0u;
            for (unsigned k0 = 0; k0 < total; k0 += 32) {
                const unsigned k = k0 + lane;
                int lo = 0, hi = 32;  // largest lane with pre <= k (all lanes run the search: shuffles are warp-wide)
#pragma unroll
                for (int step = 0; step < 5; step++) {
                    const int mid = (lo + hi) >> 1;
                    const unsigned pm = __shfl_sync(0xffffffffu, pre, mid);
                    if (pm <= k) lo = mid; else hi = mid;
                }
                const unsigned bx = __shfl_sync(0xffffffffu, rect, lo);
                const unsigned plo = __shfl_sync(0xffffffffu, pre, lo);
                SasGroupReach gr;
                gr.cx0 = __shfl_sync(0xffffffffu, my_gr.cx0, lo); gr.cx1 = __shfl_sync(0xffffffffu, my_gr.cx1, lo);
                gr.ry0 = __shfl_sync(0xffffffffu, my_gr.ry0, lo); gr.ry1 = __shfl_sync(0xffffffffu, my_gr.ry1, lo);
                if (k < total) {
                    const unsigned r = k - plo;
                    const int gx0 = bx & 0xff, gy0 = (bx >> 8) & 0xff, rw = (int)((bx >> 16) & 0xff) - gx0;
                    const int dy = (int)r / rw;
                    count_one(gr, gx0 + ((int)r - dy * rw), gy0 + dy);
                }
            }
        } else if (vis) {
            for (int y = y0; y < y1; y++)
                for (int x = x0; x < x1; x++) count_one(my_gr, x, y);
        }
    }
    if (smem_tiles) {
        __syncthreads();
        for (int k = threadIdx.x; k < n_rows * smem_tiles; k += 256) {
            const int rr = k / smem_tiles, t = k - rr * smem_tiles;
            const unsigned c = s_hist[rr * hist_stride + t];
            if (c) atomicAdd(row_tile_count + (size_t)(span.row0 + row_lo + rr) * T + t, c);
            if (tile_dyn_mask && smem_masks) {
                const unsigned m = s_hist[rr * hist_stride + smem_tiles + t];
                if (m) atomicOr(tile_dyn_mask + (size_t)(row_lo + rr) * T + t, m);
            }
        }
    }
    {  // one counter update per warp instead of one per thread
        const unsigned s = __reduce_add_sync(0xffffffffu, n_vis_thread);
        if (lane == 0 && s) atomicAdd(&counters->n_visible_src, (unsigned long long)s);
    }
}

// Projects Gaussian g for one row (camera [+ env]); writes its bin and, if visible, its record and
// radii.  Returns visibility and the tile rectangle.
template <int DEG>
__device__ __forceinline__ bool project_one_direct(const float4 *__restrict__ mean_op, const float4 *__restrict__ quat,
                                            const float4 *__restrict__ scale_link, const float *__restrict__ sh_planes, int n_pad, int g,
                                            size_t ig, const float *__restrict__ link_poses, int L, const SasCamera *__restrict__ cam, int W,
                                            int H, int tile_w, int tile_h, SasRecord *__restrict__ rec, uint2 *__restrict__ bins,
                                            uint2 *__restrict__ radii_out, const float *__restrict__ extend, int &x0, int &y0, int &x1,
                                            int &y1, float4 &reach) {

    const float4 mo = __ldg(mean_op + g);
    const float4 qq = __ldg(quat + g);
    const float4 sl = __ldg(scale_link + g);
    const int link = __float_as_int(sl.w);

    // quaternion -> rotation (gsplat _quat_to_rotmat on the normalised quaternion)
    float Rg[9];
    {
        float w = qq.x, x = qq.y, y = qq.z, z = qq.w;
        float n = sqrtf(fmaf(z, z, fmaf(y, y, fmaf(x, x, w * w))));
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
        float tx = aff3(P[0], P[1], P[2], mx, my, mz, P[9]);
        float ty = aff3(P[3], P[4], P[5], mx, my, mz, P[10]);
        float tz = aff3(P[6], P[7], P[8], mx, my, mz, P[11]);
        mx = tx; my = ty; mz = tz;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++)
                R[3 * i + j] = dot3(P[3 * i], P[3 * i + 1], P[3 * i + 2], Rg[j], Rg[3 + j], Rg[6 + j]);
    } else {
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = Rg[i];
    }

    float vm[12];
#pragma unroll
    for (int i = 0; i < 12; i++) vm[i] = cam->vm[i];
    const float cxm = aff3(vm[0], vm[1], vm[2], mx, my, mz, vm[3]);
    const float cym = aff3(vm[4], vm[5], vm[6], mx, my, mz, vm[7]);
    const float czm = aff3(vm[8], vm[9], vm[10], mx, my, mz, vm[11]);

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
                float v = dot3(Ms[3 * i], Ms[3 * i + 1], Ms[3 * i + 2], Ms[3 * j], Ms[3 * j + 1], Ms[3 * j + 2]);
                Sig[3 * i + j] = v; Sig[3 * j + i] = v;
            }
        // camera-frame covariance Rc Sigma Rc^T (Rc = rows of vm)
        float Tm[9], Cc[9];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++)
                Tm[3 * i + j] = dot3(vm[4 * i], vm[4 * i + 1], vm[4 * i + 2], Sig[j], Sig[3 + j], Sig[6 + j]);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++)
                Cc[3 * i + j] = dot3(Tm[3 * i], Tm[3 * i + 1], Tm[3 * i + 2], vm[4 * j], vm[4 * j + 1], vm[4 * j + 2]);

        // S2: EWA perspective projection (gsplat _persp_proj), blur, conic, radii, culls
        const float fx = cam->fx, fy = cam->fy, pcx = cam->cx, pcy = cam->cy;
        const float lim_x_pos = cam->lim_x_pos, lim_x_neg = cam->lim_x_neg;
        const float lim_y_pos = cam->lim_y_pos, lim_y_neg = cam->lim_y_neg;
        const float rz = 1.0f / czm;
        const float rz2 = rz * rz;
        const float txc = czm * fminf(lim_x_pos, fmaxf(-lim_x_neg, cxm * rz));
        const float tyc = czm * fminf(lim_y_pos, fmaxf(-lim_y_neg, cym * rz));
        const float ja = fx * rz, jb = -(fx * txc) * rz2, jc = fy * rz, jd = -(fy * tyc) * rz2;
        const float jc00 = fmaf(jb, Cc[6], ja * Cc[0]), jc01 = fmaf(jb, Cc[7], ja * Cc[1]), jc02 = fmaf(jb, Cc[8], ja * Cc[2]);
        const float jc11 = fmaf(jd, Cc[7], jc * Cc[4]), jc12 = fmaf(jd, Cc[8], jc * Cc[5]);
        a = fmaf(jc02, jb, jc00 * ja);
        b = fmaf(jc02, jd, jc01 * jc);
        c = fmaf(jc12, jd, jc11 * jc);
        m2x = fmaf(fx * cxm, rz, pcx);
        m2y = fmaf(fy * cym, rz, pcy);
        a = a + SAS_EPS2D; c = c + SAS_EPS2D;
        det = fmaf(a, c, -(b * b));
        vis = det > 0.0f;
        if (vis) {
            det = fmaxf(det, 1e-10f);
            const float ext = extend ? __ldg(extend + g) : SAS_EXTEND;
            rx = ceilf(ext * sqrtf(a));
            ry = ceilf(ext * sqrtf(c));
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
        float ux = dot3(P[0], P[3], P[6], dx, dy, dz);
        float uy = dot3(P[1], P[4], P[7], dx, dy, dz);
        float uz = dot3(P[2], P[5], P[8], dx, dy, dz);
        dx = ux; dy = uy; dz = uz;
    }
    float dn = sqrtf(dot3(dx, dy, dz, dx, dy, dz));
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
        for (int k = 1; k < KN; k++) acc = fmaf(bas[k], __ldg(sh_planes + (size_t)(3 * k + ch) * n_pad + g), acc);
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


// DIRECT variant (one row per CTA, nothing staged): the round-1 kernel with the round-2 arithmetic, used when a chunk of rows
// would not fit its histograms next to the staged coefficients (large images) or the span has a single row.
// One thread per (row, Gaussian) of a span: the background span runs once per VIEW (rows = distinct
// cameras), the link span once per IMAGE (rows = env x camera) -- see sas_common.cuh.  Per-(row,tile)
// counts are accumulated in shared memory per CTA (one global RED per touched tile per CTA instead
// of one per intersection); images with more tiles than fit in shared memory add to the global
// counters directly.
template <int DEG, int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS)
project_direct_kernel(const float4 *__restrict__ mean_op, const float4 *__restrict__ quat, const float4 *__restrict__ scale_link,
               const float *__restrict__ sh_planes, int n_pad, SasRowSpan span, const float *__restrict__ link_poses, int L,
               const SasCamera *__restrict__ row_cams, int W, int H, int tile_w, int tile_h, SasRecord *__restrict__ rec,
               uint2 *__restrict__ bins, uint2 *__restrict__ radii_out, unsigned int *__restrict__ row_tile_count,
               unsigned int *__restrict__ tile_dyn_mask, SasCounters *counters, int smem_tiles, int smem_masks,
               const float *__restrict__ extend) {
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
        vis = project_one_direct<DEG>(mean_op, quat, scale_link, sh_planes, n_pad, span.g0 + local,
                                      (size_t)span.rec0 + (size_t)row * span.n + local, link_poses, L, row_cams + span.row0 + row, W, H,
                                      tile_w, tile_h, rec, bins, radii_out, extend, x0, y0, x1, y1, reach);
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

cudaError_t sas_project_init() {
    cudaError_t e = cudaSuccess;
#define SAS_PROJ_ATTR(D, C, S) if (e == cudaSuccess) e = cudaFuncSetAttribute(project_kernel<D, C, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)
#define SAS_PROJ_ATTRS(D) SAS_PROJ_ATTR(D, 2, true); SAS_PROJ_ATTR(D, 3, true); SAS_PROJ_ATTR(D, 4, true)
    SAS_PROJ_ATTRS(0); SAS_PROJ_ATTRS(1); SAS_PROJ_ATTRS(2); SAS_PROJ_ATTRS(3);
#undef SAS_PROJ_ATTRS
#undef SAS_PROJ_ATTR
    return e;
}

cudaError_t sas_launch_project(const float4 *mean_op, const float4 *quat, const float4 *scale_link, const float *sh_planes,
                               int n_pad, int sh_degree, const float *link_poses, int L, const SasCamera *row_cams,
                               SasRowSpan span, int W, int H, int tile_w, int tile_h, SasRecord *rec, uint2 *bins,
                               uint2 *radii_out, unsigned int *row_tile_count, unsigned int *tile_dyn_mask, SasCounters *counters,
                               const float *extend, cudaStream_t st) {
    if (span.n <= 0 || span.rows <= 0) return cudaSuccess;
    const int T = tile_w * tile_h;
    const int smem_tiles = T <= SAS_SMEM_TILES_MAX ? T : 0;
    const int smem_masks = (tile_dyn_mask && smem_tiles && T <= SAS_SMEM_TILES_MAX / 2) ? 1 : 0;
    const int deg = sh_degree < 0 ? 0 : sh_degree > 3 ? 3 : sh_degree;
    // Rows per chunk: as many (up to PROJ_ROWS) as keep the chunk's histograms within 8 KB next to the staged
    // coefficients (4 CTAs per SM).  If not even two rows fit (large images), or the span has a single row, there is
    // nothing to reuse: one row per CTA, coefficients read straight from the global planes, no staging.
    const size_t hist_row = (size_t)smem_tiles * (1 + smem_masks) * sizeof(unsigned int);
    int rows_per_chunk = PROJ_ROWS;
    if (hist_row) rows_per_chunk = (int)((8 * 1024) / hist_row);
    if (rows_per_chunk > PROJ_ROWS) rows_per_chunk = PROJ_ROWS;
    if (rows_per_chunk > span.rows) rows_per_chunk = span.rows;
    const bool staged = rows_per_chunk >= 2;
    if (!staged) rows_per_chunk = 1;
    const size_t sh_bytes = staged ? (size_t)3 * (deg + 1) * (deg + 1) * 256 * sizeof(float) : 0;
    const size_t smem = sh_bytes + hist_row * rows_per_chunk;
    dim3 grid((span.n + 255) / 256, (span.rows + rows_per_chunk - 1) / rows_per_chunk);
    // SAS_PROJECT=n (environment, read once): minimum resident CTAs/SM the register allocation of the staged kernel
    // targets (2, 3 or 4; measured on the pusht batch: 0.115 / 0.107 / 0.103 ms)
    static int variant = -1;
    if (variant < 0) {
        const char *e = getenv("SAS_PROJECT");
        variant = e ? atoi(e) : 4;
        if (variant < 2 || variant > 4) variant = 4;
    }

#define SAS_PROJ_ARGS mean_op, quat, scale_link, sh_planes, n_pad, span, link_poses, L, row_cams, W, H, tile_w, tile_h, rec, bins, radii_out, \
                      row_tile_count, tile_dyn_mask, counters, smem_tiles, smem_masks, rows_per_chunk, extend
#define SAS_PROJ_DARGS mean_op, quat, scale_link, sh_planes, n_pad, span, link_poses, L, row_cams, W, H, tile_w, tile_h, rec, bins, radii_out, \
                       row_tile_count, tile_dyn_mask, counters, smem_tiles, smem_masks, extend
#define SAS_PROJ(D)                                                                                     \
    do {                                                                                                \
        if (!staged) project_direct_kernel<D, 4><<<dim3((span.n + 255) / 256, span.rows), 256, smem, st>>>(SAS_PROJ_DARGS); \
        else if (variant == 2) project_kernel<D, 2, true><<<grid, 256, smem, st>>>(SAS_PROJ_ARGS);      \
        else if (variant == 3) project_kernel<D, 3, true><<<grid, 256, smem, st>>>(SAS_PROJ_ARGS);      \
        else project_kernel<D, 4, true><<<grid, 256, smem, st>>>(SAS_PROJ_ARGS);                        \
    } while (0)
    switch (deg) {
        case 0: SAS_PROJ(0); break;
        case 1: SAS_PROJ(1); break;
        case 2: SAS_PROJ(2); break;
        default: SAS_PROJ(3); break;
    }
#undef SAS_PROJ
#undef SAS_PROJ_ARGS
    return cudaGetLastError();
}
