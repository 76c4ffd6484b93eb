/*
 * splat_oracle.c -- CPU ORACLE (test infrastructure, NOT a product path).
 *
 * A plain-C restatement of the per-step splat render that sim_a_splat reaches
 * through nerfstudio -> gsplat 1.5.2 (`GaussianSplat.render`,
 * /root/reference/sim_a_splat/ns_utils/nerfstudio_utils.py:123-177, the only
 * gsplat call site, :166-172) applied to a scene whose segmented link groups
 * have been rigidly moved the way `SplatHandler.draw_handler` moves them
 * (/root/reference/sim_a_splat/splat/splat_handler.py:227-314; the per-link
 * (rot,pos) come from oracle/pose_oracle.py which restates :265-281 in fp64).
 *
 * PARITY UNPINNED for the gsplat stages: gsplat 1.5.2 (pixi.lock:348), the
 * nerfstudio fork @e70d6ee (pixi.lock:387) and viser 0.2.23 (pixi.lock:479) are
 * third-party, not vendored under /root/reference and not installable here, and
 * the reference holds no golden outputs (its tests/ have no asserts).  The
 * stages below restate gsplat 1.5.2's published forward algorithm
 * (gsplat/cuda/_torch_impl.py: _quat_scale_to_covar_preci, _world_to_cam,
 * _persp_proj, _fully_fused_projection, _eval_sh_bases_fast, _isect_tiles,
 * _isect_offset_encode, _rasterize_to_pixels; gsplat/rendering.py:
 * rasterization; nerfstudio/models/splatfacto.py: get_outputs) with the pinned
 * rules of SURVEY.md section 8c.  What IS pinned against the reference itself:
 * Sigma=(RS)(RS)^T vs covariance_utils.compute_cov (tests/golden/compute_cov.npz),
 * SH2RGB/C0, the ICP factorisation and draw_handler pose algebra
 * (tests/golden/draw_handler.npz) -- see tools/make_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker / baseline.
 *
 * NUMERICS CONTRACT (shared with sim_a_splat_b200/csrc/, which is written
 * separately): every float operation is an individually rounded IEEE-754
 * binary32 operation in exactly the order written here (compile with
 * -ffp-contract=off, no -ffast-math); a fused multiply-add happens only where
 * fmaf() is written -- which since round 2 includes every 3-term inner product
 * of the projection (dot3 / aff3 below) and the SH dot product.  sqrtf and '/' are correctly rounded.  No libm
 * transcendental is used on any path that feeds an integer decision: exp() in
 * the compositor is the explicit polynomial orc_exp_neg() below.
 *
 * Pinned rules: tile 16; eps2d 0.3; near 0.01, far 1e10; radius_clip 0;
 * radii = ceil(3.33*sqrt(cov2d_xx / yy)) (no opacity shrink); alpha threshold
 * 1/255; alpha cap 0.999; stop BEFORE accumulating when T*(1-alpha) <= 1e-4;
 * sigma evaluated as fma(fma(0.5A,dx,B*dy),dx,(0.5C*dy)*dy); accumulation
 * acc = fma(colour, alpha*T, acc); pixel centres at +0.5; key = image<<(32+tile_bits) | tile<<32 | bits(depth);
 * stable sort (ties keep ascending flattened Gaussian index).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_TILE 16
#define ORC_EPS2D 0.3f
#define ORC_NEAR 0.01f
#define ORC_FAR 1e10f
#define ORC_EXTEND 3.33f
#define ORC_ALPHA_MIN (1.0f / 255.0f)
#define ORC_ALPHA_MAX 0.999f
#define ORC_T_MIN 1e-4f

typedef struct {
    int32_t N;         /* Gaussians in canonical order                          */
    int32_t sh_degree; /* 0..3, K = (d+1)^2                                     */
    int32_t L;         /* links                                                 */
    const float *means;     /* [N,3]                                            */
    const float *quats;     /* [N,4] wxyz, as stored (normalised here)          */
    const float *scales;    /* [N,3] linear (exp already applied, splat_utils.py:35-36) */
    const float *opacities; /* [N]   linear (sigmoid applied, splat_utils.py:43-45)     */
    const float *sh;        /* [N,K,3]                                          */
    const int32_t *link_id; /* [N]  -1 = background                             */
} orc_scene;

typedef struct {
    int32_t I, N, W, H, tile_w, tile_h, tile_bits, image_bits;
    int64_t M;               /* number of tile intersections                    */
    /* per (image, gaussian) */
    int32_t *radii;          /* [I,N,2]                                         */
    float *means2d;          /* [I,N,2]                                         */
    float *depths;           /* [I,N]                                           */
    float *conics;           /* [I,N,3]                                         */
    float *colors;           /* [I,N,3]  post +0.5, clamp_min 0                 */
    int32_t *tiles_per_gauss;/* [I,N]                                           */
    int32_t *tile_rect;      /* [I,N,4] min_x,min_y,max_x,max_y (exclusive max) */
    /* per intersection */
    int64_t *isect_ids;      /* [M] sorted keys                                 */
    int32_t *flatten_ids;    /* [M] sorted values = image*N + g                 */
    int32_t *isect_offsets;  /* [I,tile_h,tile_w]                               */
    /* per pixel */
    float *rgb;              /* [I,H,W,3]  acc + T*bg                           */
    float *depth;            /* [I,H,W]    expected depth acc_d/max(alpha,1e-10)*/
    float *alpha;            /* [I,H,W]    1-T                                  */
    int32_t *n_contrib;      /* [I,H,W]    Gaussians accumulated                */
    int32_t *last_idx;       /* [I,H,W]    index in sorted list of last accumulated, -1 */
    int64_t pairs_evaluated; /* sum over tiles of 256 * (Gaussians visited)     */
} orc_frame;

static int bit_length(int64_t v) { int b = 0; while (v > 0) { b++; v >>= 1; } return b; }

/* exp(x) for x <= 0: 2^(x*log2e), round-to-nearest split, degree-5 polynomial
 * evaluated with fmaf, exponent added into the float's bits. |rel err| ~2e-7. */
static inline float orc_exp_neg(float x) {
    float t = x * 1.44269502162933349609375f;
    if (t < -125.0f) t = -125.0f;
    float n = rintf(t);
    float f = t - n;
    float p = 0.001327647129073739f;
    p = fmaf(p, f, 0.009675540961325169f);
    p = fmaf(p, f, 0.05550713092088699f);
    p = fmaf(p, f, 0.24022120237350464f);
    p = fmaf(p, f, 0.6931469440460205f);
    p = fmaf(p, f, 1.0000001192092896f);
    int32_t bits; memcpy(&bits, &p, 4);
    bits += ((int32_t)n) << 23;
    float r; memcpy(&r, &bits, 4);
    return r;
}
float orc_exp_neg_export(float x) { return orc_exp_neg(x); }

/* a0 b0 + a1 b1 + a2 b2 as ONE product and TWO fused multiply-adds, in this order; aff3 adds the products onto t.
 * Every 3-term inner product of the projection is written through these two (round 2: gsplat's own CUDA contracts
 * these sums to FMAs as well; the unfused form doubled the instruction count of the projection kernel). */
static inline float dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
    return fmaf(a2, b2, fmaf(a1, b1, a0 * b0));
}
static inline float aff3(float a0, float a1, float a2, float b0, float b1, float b2, float t) {
    return fmaf(a2, b2, fmaf(a1, b1, fmaf(a0, b0, t)));
}

/* gsplat _quat_to_rotmat on F.normalize(q): row-major R[9]. */
static inline void quat_to_rotmat(const float *q, float *R) {
    float w = q[0], x = q[1], y = q[2], z = q[3];
    float n = sqrtf(fmaf(z, z, fmaf(y, y, fmaf(x, x, w * w))));
    n = fmaxf(n, 1e-12f);
    float inv = 1.0f / n;
    w = w * inv; x = x * inv; y = y * inv; z = z * inv;
    float x2 = x * x, y2 = y * y, z2 = z * z;
    float xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
    R[0] = 1.0f - 2.0f * (y2 + z2); R[1] = 2.0f * (xy - wz);        R[2] = 2.0f * (xz + wy);
    R[3] = 2.0f * (xy + wz);        R[4] = 1.0f - 2.0f * (x2 + z2); R[5] = 2.0f * (yz - wx);
    R[6] = 2.0f * (xz - wy);        R[7] = 2.0f * (yz + wx);        R[8] = 1.0f - 2.0f * (x2 + y2);
}

/* C = A*B, 3x3 row-major, each entry dot3(row of A, column of B). */
static inline void mat3_mul(const float *A, const float *B, float *C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            C[3 * i + j] = dot3(A[3 * i], A[3 * i + 1], A[3 * i + 2], B[j], B[3 + j], B[6 + j]);
}
/* C = A*B^T */
static inline void mat3_mul_bt(const float *A, const float *B, float *C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            C[3 * i + j] = dot3(A[3 * i], A[3 * i + 1], A[3 * i + 2], B[3 * j], B[3 * j + 1], B[3 * j + 2]);
}

/* gsplat _eval_sh_bases_fast (Sloan), K=(deg+1)^2 basis values for unit dir. */
static inline void sh_bases(int deg, float x, float y, float z, float *b) {
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

/* S0-S4 for one (image, gaussian). viewmat row-major 4x4 world->cam (OpenCV
 * axes, +z forward), K row-major 3x3, campos = camera origin in world (passed
 * in so both sides use the same fp32 value), pose = [rot(9) | pos(3)] or NULL. */
/* Radius rule.  0 (pinned default): gsplat's torch implementation, radii = ceil(3.33 sqrt(cov2d)).
 * 1 ("cuda_1_5"): gsplat 1.5's CUDA projection also shrinks the extend with the opacity -- a Gaussian whose peak alpha
 * is o reaches alpha >= 1/255 only within sqrt(2 ln(255 o)) sigma -- extend = min(3.33, sqrt(2 ln(o / (1/255)))) and
 * culls o < 1/255 [RECALLED].  The factor depends on the Gaussian alone; it is evaluated once per Gaussian in double
 * precision and rounded to float (the product does the same on the host at scene upload, so both sides multiply by
 * the same float).  Images are unaffected: a pixel outside either box has alpha < 1/255. */
static float orc_extend(int radius_rule, float opacity) {
    if (radius_rule != 1) return ORC_EXTEND;
    if (opacity < ORC_ALPHA_MIN) return 0.0f;
    double e = sqrt(2.0 * log((double)opacity / (double)ORC_ALPHA_MIN));
    float ef = (float)e;
    return ef < ORC_EXTEND ? ef : ORC_EXTEND;
}
float orc_extend_export(int radius_rule, float opacity) { return orc_extend(radius_rule, opacity); }

static void project_one(const orc_scene *s, int g, float extend, const float *pose, const float *vm, const float *Kc,
                        const float *campos, int W, int H, int tile_w, int tile_h,
                        int32_t *radii, float *mean2d, float *depth, float *conic, float *color,
                        int32_t *ntiles, int32_t *rect) {
    radii[0] = radii[1] = 0; mean2d[0] = mean2d[1] = 0.0f; *depth = 0.0f;
    conic[0] = conic[1] = conic[2] = 0.0f; color[0] = color[1] = color[2] = 0.0f;
    *ntiles = 0; rect[0] = rect[1] = rect[2] = rect[3] = 0;

    const float *mu = s->means + 3 * g;
    float R[9], Rg[9];
    quat_to_rotmat(s->quats + 4 * g, Rg);
    float mx = mu[0], my = mu[1], mz = mu[2];
    if (pose) { /* S0: mu' = rot*mu + pos ; R' = rot*R(q)  (splat_handler.py:283-288 semantics) */
        float tx = aff3(pose[0], pose[1], pose[2], mx, my, mz, pose[9]);
        float ty = aff3(pose[3], pose[4], pose[5], mx, my, mz, pose[10]);
        float tz = aff3(pose[6], pose[7], pose[8], mx, my, mz, pose[11]);
        mx = tx; my = ty; mz = tz;
        mat3_mul(pose, Rg, R);
    } else {
        memcpy(R, Rg, sizeof(R));
    }
    /* world -> camera mean */
    float cx_ = aff3(vm[0], vm[1], vm[2], mx, my, mz, vm[3]);
    float cy_ = aff3(vm[4], vm[5], vm[6], mx, my, mz, vm[7]);
    float cz_ = aff3(vm[8], vm[9], vm[10], mx, my, mz, vm[11]);
    if (cz_ < ORC_NEAR || cz_ > ORC_FAR) return;

    /* S1: Sigma = (R S)(R S)^T */
    const float *sc = s->scales + 3 * g;
    float Ms[9], Sig[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Ms[3 * i + j] = R[3 * i + j] * sc[j];
    mat3_mul_bt(Ms, Ms, Sig);
    /* camera-frame covariance Rc Sigma Rc^T */
    float Rc[9] = { vm[0], vm[1], vm[2], vm[4], vm[5], vm[6], vm[8], vm[9], vm[10] };
    float Tm[9], Cc[9];
    mat3_mul(Rc, Sig, Tm);
    mat3_mul_bt(Tm, Rc, Cc);

    /* S2: EWA perspective projection (gsplat _persp_proj) */
    float fx = Kc[0], fy = Kc[4], pcx = Kc[2], pcy = Kc[5];
    float tan_fovx = 0.5f * (float)W / fx, tan_fovy = 0.5f * (float)H / fy;
    float lim_x_pos = ((float)W - pcx) / fx + 0.3f * tan_fovx;
    float lim_x_neg = pcx / fx + 0.3f * tan_fovx;
    float lim_y_pos = ((float)H - pcy) / fy + 0.3f * tan_fovy;
    float lim_y_neg = pcy / fy + 0.3f * tan_fovy;
    float rz = 1.0f / cz_;
    float rz2 = rz * rz;
    float txc = cz_ * fminf(lim_x_pos, fmaxf(-lim_x_neg, cx_ * rz));
    float tyc = cz_ * fminf(lim_y_pos, fmaxf(-lim_y_neg, cy_ * rz));
    float ja = fx * rz, jb = -(fx * txc) * rz2, jc = fy * rz, jd = -(fy * tyc) * rz2;
    float jc00 = fmaf(jb, Cc[6], ja * Cc[0]), jc01 = fmaf(jb, Cc[7], ja * Cc[1]), jc02 = fmaf(jb, Cc[8], ja * Cc[2]);
    float jc11 = fmaf(jd, Cc[7], jc * Cc[4]), jc12 = fmaf(jd, Cc[8], jc * Cc[5]);
    float a = fmaf(jc02, jb, jc00 * ja);
    float b = fmaf(jc02, jd, jc01 * jc);
    float c = fmaf(jc12, jd, jc11 * jc);
    float m2x = fmaf(fx * cx_, rz, pcx);
    float m2y = fmaf(fy * cy_, rz, pcy);
    a = a + ORC_EPS2D; c = c + ORC_EPS2D;
    float det = fmaf(a, c, -(b * b));
    if (!(det > 0.0f)) return;
    det = fmaxf(det, 1e-10f);
    float rx = ceilf(extend * sqrtf(a));
    float ry = ceilf(extend * sqrtf(c));
    if (rx <= 0.0f && ry <= 0.0f) return;
    if (m2x + rx <= 0.0f || m2x - rx >= (float)W || m2y + ry <= 0.0f || m2y - ry >= (float)H) return;

    /* saturating conversion (a screen-filling Gaussian a few millimetres from the camera can exceed int32) */
    radii[0] = (int32_t)fminf(rx, 2147483520.0f); radii[1] = (int32_t)fminf(ry, 2147483520.0f);
    mean2d[0] = m2x; mean2d[1] = m2y; *depth = cz_;
    conic[0] = c / det; conic[1] = -b / det; conic[2] = a / det;

    /* S3: SH colour; view direction in the link's rest frame: rot^T (mu' - campos) */
    float dx = mx - campos[0], dy = my - campos[1], dz = mz - campos[2];
    if (pose) {
        float ux = dot3(pose[0], pose[3], pose[6], dx, dy, dz);
        float uy = dot3(pose[1], pose[4], pose[7], dx, dy, dz);
        float uz = dot3(pose[2], pose[5], pose[8], dx, dy, dz);
        dx = ux; dy = uy; dz = uz;
    }
    float dn = sqrtf(dot3(dx, dy, dz, dx, dy, dz));
    dn = fmaxf(dn, 1e-12f);
    float dinv = 1.0f / dn;
    dx = dx * dinv; dy = dy * dinv; dz = dz * dinv;
    int Kn = (s->sh_degree + 1) * (s->sh_degree + 1);
    float bas[16];
    sh_bases(s->sh_degree, dx, dy, dz, bas);
    const float *co = s->sh + (size_t)g * Kn * 3;
    for (int ch = 0; ch < 3; ch++) {
        float acc = bas[0] * co[ch];
        for (int k = 1; k < Kn; k++) acc = fmaf(bas[k], co[3 * k + ch], acc);
        color[ch] = fmaxf(acc + 0.5f, 0.0f);
    }

    /* S4: tile rectangle (gsplat _isect_tiles) */
    float tmx = m2x / (float)ORC_TILE, tmy = m2y / (float)ORC_TILE;
    float trx = rx / (float)ORC_TILE, try_ = ry / (float)ORC_TILE;
    int x0 = (int)floorf(tmx - trx), y0 = (int)floorf(tmy - try_);
    int x1 = (int)ceilf(tmx + trx), y1 = (int)ceilf(tmy + try_);
    x0 = x0 < 0 ? 0 : (x0 > tile_w ? tile_w : x0); x1 = x1 < 0 ? 0 : (x1 > tile_w ? tile_w : x1);
    y0 = y0 < 0 ? 0 : (y0 > tile_h ? tile_h : y0); y1 = y1 < 0 ? 0 : (y1 > tile_h ? tile_h : y1);
    rect[0] = x0; rect[1] = y0; rect[2] = x1; rect[3] = y1;
    *ntiles = (x1 - x0) * (y1 - y0);
}

/* stable LSD radix sort of (key,val) on bits [0,nbits) */
static void radix_sort_pairs(int64_t *keys, int32_t *vals, int64_t n, int nbits) {
    if (n <= 1) return;
    int64_t *k2 = (int64_t *)malloc(sizeof(int64_t) * n);
    int32_t *v2 = (int32_t *)malloc(sizeof(int32_t) * n);
    int64_t *ka = keys, *kb = k2; int32_t *va = vals, *vb = v2;
    for (int shift = 0; shift < nbits; shift += 8) {
        int64_t cnt[257]; memset(cnt, 0, sizeof(cnt));
        for (int64_t i = 0; i < n; i++) cnt[((ka[i] >> shift) & 0xff) + 1]++;
        for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
        for (int64_t i = 0; i < n; i++) { int64_t p = cnt[(ka[i] >> shift) & 0xff]++; kb[p] = ka[i]; vb[p] = va[i]; }
        int64_t *tk = ka; ka = kb; kb = tk; int32_t *tv = va; va = vb; vb = tv;
    }
    if (ka != keys) { memcpy(keys, ka, sizeof(int64_t) * n); memcpy(vals, va, sizeof(int32_t) * n); }
    free(k2); free(v2);
}

void orc_frame_free(orc_frame *f) {
    if (!f) return;
    free(f->radii); free(f->means2d); free(f->depths); free(f->conics); free(f->colors);
    free(f->tiles_per_gauss); free(f->tile_rect); free(f->isect_ids); free(f->flatten_ids);
    free(f->isect_offsets); free(f->rgb); free(f->depth); free(f->alpha); free(f->n_contrib); free(f->last_idx);
    free(f);
}

/*
 * Full frame batch. link_poses [E,L,12] (rot row-major | pos) may be NULL when
 * L == 0; viewmats [I,16], Ks [I,9], campos [I,3], env_of_image [I].
 * stages: 1 = projection only (S0-S4), 2 = + emit/sort/offsets (S5-S7),
 * 3 = + compositing (S8).  keep_dense = 0 skips nothing (all intermediates are
 * always kept; the caller frees with orc_frame_free).
 */
orc_frame *orc_render2(const orc_scene *s, const float *link_poses, int E, const float *viewmats,
                       const float *Ks, const float *campos, const int32_t *env_of_image, int I,
                       int W, int H, const float *bg, int stages, int radius_rule);
orc_frame *orc_render(const orc_scene *s, const float *link_poses, int E, const float *viewmats,
                      const float *Ks, const float *campos, const int32_t *env_of_image, int I,
                      int W, int H, const float *bg, int stages) {
    return orc_render2(s, link_poses, E, viewmats, Ks, campos, env_of_image, I, W, H, bg, stages, 0);
}
orc_frame *orc_render2(const orc_scene *s, const float *link_poses, int E, const float *viewmats,
                       const float *Ks, const float *campos, const int32_t *env_of_image, int I,
                       int W, int H, const float *bg, int stages, int radius_rule) {
    orc_frame *f = (orc_frame *)calloc(1, sizeof(orc_frame));
    int N = s->N;
    int tile_w = (W + ORC_TILE - 1) / ORC_TILE, tile_h = (H + ORC_TILE - 1) / ORC_TILE;
    f->I = I; f->N = N; f->W = W; f->H = H; f->tile_w = tile_w; f->tile_h = tile_h;
    f->tile_bits = bit_length((int64_t)tile_w * tile_h);
    f->image_bits = bit_length(I);
    size_t IN = (size_t)I * N;
    f->radii = (int32_t *)malloc(sizeof(int32_t) * 2 * IN);
    f->means2d = (float *)malloc(sizeof(float) * 2 * IN);
    f->depths = (float *)malloc(sizeof(float) * IN);
    f->conics = (float *)malloc(sizeof(float) * 3 * IN);
    f->colors = (float *)malloc(sizeof(float) * 3 * IN);
    f->tiles_per_gauss = (int32_t *)malloc(sizeof(int32_t) * IN);
    f->tile_rect = (int32_t *)malloc(sizeof(int32_t) * 4 * IN);
    (void)E;

#pragma omp parallel for schedule(static) collapse(2)
    for (int im = 0; im < I; im++) {
        for (int g = 0; g < N; g++) {
            size_t ig = (size_t)im * N + g;
            int l = s->link_id ? s->link_id[g] : -1;
            const float *pose = (l >= 0 && link_poses) ? link_poses + ((size_t)env_of_image[im] * s->L + l) * 12 : NULL;
            project_one(s, g, orc_extend(radius_rule, s->opacities[g]), pose, viewmats + 16 * im, Ks + 9 * im, campos + 3 * im, W, H, tile_w, tile_h,
                        f->radii + 2 * ig, f->means2d + 2 * ig, f->depths + ig, f->conics + 3 * ig,
                        f->colors + 3 * ig, f->tiles_per_gauss + ig, f->tile_rect + 4 * ig);
        }
    }
    if (stages < 2) return f;

    /* S4 scan + S5 emit: image-major, Gaussian-minor, rows then columns */
    int64_t *img_start = (int64_t *)malloc(sizeof(int64_t) * (I + 1));
    img_start[0] = 0;
    for (int im = 0; im < I; im++) {
        int64_t c = 0;
        for (int g = 0; g < N; g++) c += f->tiles_per_gauss[(size_t)im * N + g];
        img_start[im + 1] = img_start[im] + c;
    }
    int64_t M = img_start[I];
    f->M = M;
    f->isect_ids = (int64_t *)malloc(sizeof(int64_t) * (M > 0 ? M : 1));
    f->flatten_ids = (int32_t *)malloc(sizeof(int32_t) * (M > 0 ? M : 1));
    int tile_bits = f->tile_bits;
#pragma omp parallel for schedule(dynamic, 1)
    for (int im = 0; im < I; im++) {
        int64_t p = img_start[im];
        for (int g = 0; g < N; g++) {
            size_t ig = (size_t)im * N + g;
            if (f->tiles_per_gauss[ig] == 0) continue;
            const int32_t *r = f->tile_rect + 4 * ig;
            uint32_t dbits; memcpy(&dbits, f->depths + ig, 4);
            for (int y = r[1]; y < r[3]; y++)
                for (int x = r[0]; x < r[2]; x++) {
                    int64_t tile = (int64_t)y * tile_w + x;
                    f->isect_ids[p] = ((int64_t)im << (32 + tile_bits)) | (tile << 32) | (int64_t)dbits;
                    f->flatten_ids[p] = (int32_t)ig;
                    p++;
                }
        }
        /* S6: stable sort. Emission is image-major, so sorting each image's
         * segment on the low 32+tile_bits bits equals the global stable sort
         * on [0, 32+tile_bits+image_bits). */
        radix_sort_pairs(f->isect_ids + img_start[im], f->flatten_ids + img_start[im],
                         img_start[im + 1] - img_start[im], 32 + tile_bits);
    }
    /* S7: first index of each (image,tile) run (gsplat _isect_offset_encode) */
    int64_t n_tiles = (int64_t)tile_w * tile_h;
    f->isect_offsets = (int32_t *)malloc(sizeof(int32_t) * I * n_tiles);
    {
        int64_t p = 0;
        for (int64_t t = 0; t < (int64_t)I * n_tiles; t++) {
            int64_t im = t / n_tiles, tile = t % n_tiles;
            int64_t want = (im << tile_bits) | tile;
            while (p < M && (f->isect_ids[p] >> 32) < want) p++;
            f->isect_offsets[t] = (int32_t)p;
        }
    }
    free(img_start);
    if (stages < 3) return f;

    /* S8: per-pixel front-to-back compositing (gsplat _rasterize_to_pixels, RGB+ED) */
    size_t P = (size_t)H * W;
    f->rgb = (float *)malloc(sizeof(float) * 3 * I * P);
    f->depth = (float *)malloc(sizeof(float) * I * P);
    f->alpha = (float *)malloc(sizeof(float) * I * P);
    f->n_contrib = (int32_t *)malloc(sizeof(int32_t) * I * P);
    f->last_idx = (int32_t *)malloc(sizeof(int32_t) * I * P);
    int64_t pairs = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : pairs)
    for (int64_t t = 0; t < (int64_t)I * n_tiles; t++) {
        int im = (int)(t / n_tiles); int tile = (int)(t % n_tiles);
        int ty = tile / tile_w, tx = tile % tile_w;
        int64_t lo = f->isect_offsets[t];
        int64_t hi = (t + 1 < (int64_t)I * n_tiles) ? f->isect_offsets[t + 1] : M;
        int64_t deepest = 0;
        for (int py = ty * ORC_TILE; py < ty * ORC_TILE + ORC_TILE && py < H; py++)
            for (int px = tx * ORC_TILE; px < tx * ORC_TILE + ORC_TILE && px < W; px++) {
                float pxf = (float)px + 0.5f, pyf = (float)py + 0.5f;
                float T = 1.0f, ar = 0.0f, ag = 0.0f, ab = 0.0f, ad = 0.0f;
                int32_t nc = 0, last = -1;
                int64_t i;
                for (i = lo; i < hi; i++) {
                    size_t ig = (size_t)f->flatten_ids[i];
                    int g = (int)(ig % (size_t)N);
                    const float *m2 = f->means2d + 2 * ig; const float *cn = f->conics + 3 * ig;
                    float dx = m2[0] - pxf, dy = m2[1] - pyf;
                    /* sigma = 0.5*(A dx^2 + C dy^2) + B dx dy, pinned in this
                     * FMA form: dx*(0.5A*dx + B*dy) + (0.5C*dy)*dy */
                    float hA = 0.5f * cn[0], hC = 0.5f * cn[2];
                    float sigma = fmaf(fmaf(hA, dx, cn[1] * dy), dx, (hC * dy) * dy);
                    if (sigma < 0.0f) continue;
                    float al = fminf(ORC_ALPHA_MAX, s->opacities[g] * orc_exp_neg(-sigma));
                    if (al < ORC_ALPHA_MIN) continue;
                    float nT = T * (1.0f - al);
                    if (nT <= ORC_T_MIN) { i++; break; }
                    float vis = al * T;
                    const float *col = f->colors + 3 * ig;
                    ar = fmaf(col[0], vis, ar); ag = fmaf(col[1], vis, ag); ab = fmaf(col[2], vis, ab);
                    ad = fmaf(f->depths[ig], vis, ad);
                    T = nT; nc++; last = (int32_t)i;
                }
                if (i - lo > deepest) deepest = i - lo;
                size_t pi = (size_t)im * P + (size_t)py * W + px;
                float A = 1.0f - T;
                f->rgb[3 * pi + 0] = fmaf(T, bg[0], ar);
                f->rgb[3 * pi + 1] = fmaf(T, bg[1], ag);
                f->rgb[3 * pi + 2] = fmaf(T, bg[2], ab);
                f->alpha[pi] = A;
                f->depth[pi] = ad / fmaxf(A, 1e-10f);
                f->n_contrib[pi] = nc; f->last_idx[pi] = last;
            }
        pairs += 256 * deepest;
    }
    f->pairs_evaluated = pairs;
    return f;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* Sigma = (R S)(R S)^T for N Gaussians -> [N,9]; used to pin against
 * /root/reference/sim_a_splat/ellipsoids/covariance_utils.py:152-157. */
void orc_compute_cov(const float *quats, const float *scales, int N, float *out) {
    for (int g = 0; g < N; g++) {
        float R[9], Ms[9];
        quat_to_rotmat(quats + 4 * g, R);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Ms[3 * i + j] = R[3 * i + j] * scales[3 * g + j];
        mat3_mul_bt(Ms, Ms, out + 9 * g);
    }
}
