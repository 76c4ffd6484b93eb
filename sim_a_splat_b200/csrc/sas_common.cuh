// sas_common.cuh -- shared device structs and helpers of libsas_b200 (sm_100a).
//
// Numerics contract (see DESIGN.md "Numerics"): this library is compiled with
// -fmad=false, so every '*' '+' '-' '/' below is one individually rounded IEEE
// binary32 operation in the order written; a fused multiply-add happens only
// where fmaf() is written.  sqrtf and '/' are IEEE (nvcc defaults, no
// --use_fast_math).  No hardware approximation (ex2.approx, rsqrt, rcp) is used
// anywhere that feeds an integer decision (radii, tile ids, depth keys, the
// alpha / transmittance tests).
//
// Work decomposition (DESIGN.md "Views and buckets"): images that share a camera
// (same viewmat and K, e.g. the fixed camera of 256 parallel envs) share one
// VIEW.  Background Gaussians (link_id < 0, the first n_bg in canonical order)
// are projected, binned and depth-sorted once per view; link Gaussians once per
// image; the two sorted lists of every (image, tile) are merged into the exact
// order a global stable sort of gsplat's (image|tile|depth) keys would give.
// "Rows" below are the C view rows followed by the I image rows.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SAS_TILE_PX 16
#define SAS_EPS2D 0.3f
#define SAS_NEAR 0.01f
#define SAS_FAR 1e10f
#define SAS_EXTEND 3.33f
#define SAS_ALPHA_MIN (1.0f / 255.0f)
#define SAS_ALPHA_MAX 0.999f
#define SAS_T_MIN 1e-4f
// sigma above which alpha < 1/255 for every opacity <= 1: exp(-5.55) = 3.887e-3 < 1/255 = 3.9216e-3
// (0.9 % margin against the 2e-7 relative error of sas_exp_neg).  Pure work-skipping: the exact
// alpha test below it is unchanged.
#define SAS_SIGMA_SKIP 5.55f

// Per-row camera block, filled on the host by sas_set_cameras.
struct SasCamera {
    float vm[12];     // world->camera rows 0..2 of the 4x4 (row-major): R | t
    float fx, fy, cx, cy;
    float campos[3];  // camera origin in world, -R^T t computed in fp64 on the host
    int32_t env;      // env whose link poses this image uses (image rows only)
    // gsplat _persp_proj clamp limits, computed once per camera on the host with the same
    // individually rounded fp32 operations the oracle performs per Gaussian
    float lim_x_pos, lim_x_neg, lim_y_pos, lim_y_neg;
};

// Projected Gaussian, 48 B, written per visible (row, gaussian).
struct __align__(16) SasRecord {
    float4 a;  // m2x, m2y, 0.5*conic_a, conic_b
    float4 b;  // 0.5*conic_c, opacity, radius_x, radius_y
    float4 c;  // r, g, b, depth
};

// exp(x), x <= 0.  2^(x*log2e) with a round-to-nearest split and a degree-5
// polynomial in explicit fmaf; the exponent is added into the float's bits.
// Bit-identical to oracle/splat_oracle.c:orc_exp_neg (same operations, same
// order); |rel err| ~ 2e-7.  ex2.approx is not reproducible off-device, and the
// transmittance test T*(1-alpha) <= 1e-4 must flip on the same Gaussian in both.
__device__ __forceinline__ float sas_exp_neg(float x) {
    float t = x * 1.44269502162933349609375f;
    t = fmaxf(t, -125.0f);
    // rintf(t) as (t + 1.5*2^23) - 1.5*2^23: the same round-to-nearest-even integer for |t| < 2^22, on the
    // FMA pipe instead of the conversion unit (FRND + F2I); the low bits of the biased sum ARE the integer,
    // so (bits << 23) is n << 23 (the bias 0x4B400000 has nine zero low bits and shifts out).
    const float nf = t + 12582912.0f;
    const float n = nf - 12582912.0f;
    float f = t - n;
    float p = 0.001327647129073739f;
    p = fmaf(p, f, 0.009675540961325169f);
    p = fmaf(p, f, 0.05550713092088699f);
    p = fmaf(p, f, 0.24022120237350464f);
    p = fmaf(p, f, 0.6931469440460205f);
    p = fmaf(p, f, 1.0000001192092896f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(nf) << 23));
}

// The same function for -8 <= x <= 0 (the compositor only evaluates sigma <= SAS_SIGMA_SKIP): the clamp at -125
// cannot bind there, so it is dropped -- every remaining operation, and therefore every result bit, is unchanged.
__device__ __forceinline__ float sas_exp_neg_small(float x) {
    const float t = x * 1.44269502162933349609375f;
    const float nf = t + 12582912.0f;
    const float n = nf - 12582912.0f;
    const float f = t - n;
    float p = 0.001327647129073739f;
    p = fmaf(p, f, 0.009675540961325169f);
    p = fmaf(p, f, 0.05550713092088699f);
    p = fmaf(p, f, 0.24022120237350464f);
    p = fmaf(p, f, 0.6931469440460205f);
    p = fmaf(p, f, 1.0000001192092896f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(nf) << 23));
}

__device__ __forceinline__ void sas_cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void sas_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void sas_cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Device-side counters of one render call.
struct SasCounters {
    unsigned long long n_isects_src;   // bucket entries emitted (background per view + links per image)
    unsigned long long n_isects;       // merged per-image intersections (= gsplat's n_isects)
    unsigned long long n_visible_src;  // projected records written
    unsigned long long n_staged;       // records the compositor staged into shared memory
    unsigned int max_tile_list;        // longest merged (image,tile) list
    unsigned int overflow;             // bit 0: source pairs exceeded capacity, bit 1: merged ids exceeded capacity
};

// What one projection / emission launch covers: a contiguous Gaussian range for a set of rows.
struct SasRowSpan {
    int g0;          // first canonical Gaussian index of the range
    int n;           // Gaussians in the range
    int rows;        // number of rows (views for the background span, images for the link span)
    int row0;        // index of the span's first row among all rows (0 for views, C for images)
    long long rec0;  // offset of the span's first record / bin in the rec / bins arrays
};

// Bucket-sort size classes (entries): threads / shared memory differ per class (sas_sort.cu).
#define SAS_SORT_CLASSES 5
#define SAS_SORT_CAP0 512
#define SAS_SORT_CAP1 2048
#define SAS_SORT_CAP2 5632
#define SAS_SORT_CAP3 12288
// buckets beyond shared memory: more than SAS_BIG_MIN_BUCKET and at most SAS_BIG_MAX_BUCKET entries go through the
// multi-CTA chunk sort + merge passes (when enabled), the others through the one-CTA global-memory kernel (sas_sort.cu)
#define SAS_BIG_MIN_BUCKET (4ll * SAS_SORT_CAP3)
#define SAS_BIG_MAX_BUCKET (64ll * SAS_SORT_CAP3)
#define SAS_COMP_CLASSES 32
// Images with at most this many tiles count / reserve bucket slots through per-CTA shared-memory
// histograms (projection and emission); larger images fall back to global atomics per intersection.
#define SAS_SMEM_TILES_MAX 8192

// Device-side work lists of one render call (sas_work.cu); zeroed with the counters.
struct SasWork {
    unsigned sort_count[SAS_SORT_CLASSES], sort_start[SAS_SORT_CLASSES], sort_fill[SAS_SORT_CLASSES], sort_take[SAS_SORT_CLASSES];
    unsigned comp_count[SAS_COMP_CLASSES], comp_start[SAS_COMP_CLASSES], comp_fill[SAS_COMP_CLASSES];
    unsigned comp_total, comp_take;
    // buckets beyond shared memory (sas_sort.cu): (bucket, chunk) items and one take counter per kernel of the path
    // ([0] chunk sort, [p] merge pass p)
    unsigned big_items, big_take[8];
};

// Sub-tile granularity of the static shortcut: a 16x16 tile is 4 columns (4 px) x 8 rows (2 px) of
// 4x2-pixel GROUPS.  The compositor gives each warp an 8x4 footprint (warp = (ry>>1)*2 + (cx>>1)) and
// each of its four 8-lane groups one 4x2 block (group = (ry&1)*2 + (cx&1)); bit warp*4+group of a
// tile's 32-bit mask says "some link Gaussian of this image can reach that block" (its culling
// extents intersect it).  Blocks no link Gaussian can reach are pixel-identical to the view's
// background-only tile and are copied instead of composited.
__host__ __device__ __forceinline__ int sas_group_bit(int cx /*0..3*/, int ry /*0..7*/) {
    return ((ry >> 1) * 2 + (cx >> 1)) * 4 + (ry & 1) * 2 + (cx & 1);
}

// groups of tile t that have at least one pixel inside the image (ragged right / bottom edge)
__device__ __forceinline__ unsigned sas_inside_group_mask(int t, int tile_w, int W, int H) {
    const int ty = t / tile_w, tx = t - ty * tile_w;
    if ((tx + 1) * SAS_TILE_PX <= W && (ty + 1) * SAS_TILE_PX <= H) return 0xffffffffu;
    unsigned m = 0u;
    for (int ry = 0; ry < 8; ry++)
        for (int cx = 0; cx < 4; cx++)
            if (tx * SAS_TILE_PX + 4 * cx < W && ty * SAS_TILE_PX + 2 * ry < H) m |= 1u << sas_group_bit(cx, ry);
    return m;
}

// Which 4x2-pixel groups a Gaussian's culling extents reach, as inclusive ranges of GLOBAL group
// columns (4 px wide) and rows (2 px high): column c covers pixel centres 4c+0.5 .. 4c+3.5, so it is
// reached iff xh >= 4c+0.5 and xl <= 4c+3.5.  (A rounding ulp at an exact boundary is harmless: the
// extents carry a margin of 0.02 px + 0.1 %, see sas_project.cu project_one.)
struct SasGroupReach { int cx0, cx1, ry0, ry1; };
__device__ __forceinline__ SasGroupReach sas_group_reach(const float4 reach) {
    SasGroupReach g;
    g.cx0 = (int)ceilf((reach.x - 3.5f) * 0.25f);
    g.cx1 = (int)floorf((reach.y - 0.5f) * 0.25f);
    g.ry0 = (int)ceilf((reach.z - 1.5f) * 0.5f);
    g.ry1 = (int)floorf((reach.w - 0.5f) * 0.5f);
    return g;
}
// ... restricted to tile (tx, ty): 32-bit mask in the layout of sas_common.cuh sas_group_bit.
__device__ __forceinline__ unsigned sas_group_reach_mask(const SasGroupReach g, int tx, int ty) {
    const int c0 = max(g.cx0 - 4 * tx, 0), c1 = min(g.cx1 - 4 * tx, 3);
    const int r0 = max(g.ry0 - 8 * ty, 0), r1 = min(g.ry1 - 8 * ty, 7);
    if (c0 > c1 || r0 > r1) return 0u;
    const unsigned xm = ((2u << (c1 - c0)) - 1u) << c0;  // 4 bits
    const unsigned ym = ((2u << (r1 - r0)) - 1u) << r0;  // 8 bits
    const unsigned xs = (xm & 3u) | ((xm & 12u) << 2);   // columns 0,1 -> bits 0,1; columns 2,3 -> bits 4,5
    // row r -> bit offset (r>>1)*8 + (r&1)*2; the shifted copies of xs never overlap, so multiply = OR of shifts
    const unsigned e = (ym & 1u) | ((ym & 2u) << 1) | ((ym & 4u) << 6) | ((ym & 8u) << 7) | ((ym & 16u) << 12) | ((ym & 32u) << 13) |
                       ((ym & 64u) << 18) | ((ym & 128u) << 19);
    return xs * e;
}

// Experiment switches, environment variable SAS_VARIANT (bit mask, read once; default 0x32cf):
//   bit 0  compositor: per-group (4x2) hit masks instead of one per warp footprint
//   bit 1  group-level static shortcut (off: every group of a tile with link Gaussians is composited)
//   bit 2  bucket-sort size classes run on concurrent streams (fork / join on the caller's stream)
//   bit 3  merge skips tiles whose link bucket is empty (their ids are produced on demand for debug fetches)
//   bit 6  compositor: each staged record's 32-bit group reach mask is computed once per CTA (by the thread
//          that staged it) instead of every warp testing every record against its own four groups
//   bit 7  compositor: warp -> footprint map rotated with the work-item index (evens out the four schedulers)
//   bit 8  scans + work lists as eleven small kernels even when the single-CTA fused kernel applies
//   bit 9  compositor: the groups link Gaussians can reach packed four per warp, a tile = one or two CTAs of 4 warps
//   bit 10 compositor (with bit 9): "split" -- alpha evaluated record-parallel (lane = hit, 8 pixels each), only the
//          transmittance chain serial per pixel (sas_composite.cu MODE 3)
//   bit 11 compositor (with bit 9): bulk-async staging -- one 48-byte cp.async.bulk per record completing on an mbarrier
//          with expect_tx (UBLKCP / SYNCS in SASS) instead of three 16-byte cp.async (LDGSTS) + wait_group
//   bit 12 merge: the lists handed to the compositor leave out the background records that reach none of the groups the
//          tile's link Gaussians can reach (bg_reach_kernel + compaction in merge_kernel); only when images share views
//   bit 13 (with bit 12) the view buckets' records are materialised in depth order with EXACT reach masks (bg_reach_kernel);
//          the merged lists name background records by their position there, the compositor stages them from there and
//          takes their masks as they are
//   bit 14 (with bits 12 + 13; ablation, off) a tile composited by two CTAs gets TWO lists (merge_kernel<true>): a background
//          entry goes to the list of the CTA(s) whose groups its exact mask touches, a link entry to both -- instead of both
//          CTAs staging the whole list.  Measured +1 % on the pusht batch, -1.4 % on the configs[3] shard
//   bits 4-5  compositor CTAs per tile: 0 = one (8 warps), 1 = two (4 warps each), 2-3 = four (2 warps each)
int sas_variant_flags();

// Outputs of the compositor: per-image buffers (any may be null) ...
struct SasOutputs {
    float *rgb, *depth, *alpha;
    uint8_t *rgb_u8;
    int *n_contrib;
    uint16_t *depth_f16;  // IEEE binary16 bits of depth
};
// ... and the per-view buffers holding the background-only tiles shared by the images of a view.
struct SasViewBuf {
    float *acc;      // [C,H,W,4] accumulated r,g,b and the transmittance T
    float *depth;    // [C,H,W]   expected depth
    int *n_contrib;  // [C,H,W]
};

// Constants of the link-pose algebra (sas_pose.cu); fk_R / fk_t / offset are device pointers [L,9] / [L,3] / [L,3].
struct SasPoseModel {
    int L;
    double s, Ri[9], ti[3];
    const double *fk_R, *fk_t, *offset;
};

// URDF joint tree for the device forward kinematics (sas_pose.cu fk_kernel); all pointers are device pointers.
struct SasFkModel {
    int J, B;                 // joints (parent-before-child order), bodies produced
    const int *parent;        // [J] link frame the joint hangs from (0 = root link, j + 1 = child frame of joint j)
    const int *type;          // [J] 0 fixed, 1 revolute / continuous, 2 prismatic
    const int *qindex;        // [J] column of q that drives the joint (a mimic joint: its source's column), -1 = none
    const int *body_frame;    // [B] link frame of output body b
    const double *mimic;      // [J,2] multiplier, offset (1, 0 for ordinary joints)
    const double *origin;     // [J,12] joint origin: rotation row-major (9) | translation (3)
    const double *axis;       // [J,3] unit axis
    double base[12];          // world pose of the root link
};
cudaError_t sas_launch_fk(const double *q, int A, const SasFkModel &m, int E, const double *extra, int L, double *body, cudaStream_t st);
int sas_fk_max_joints();

// Launchers (each returns the cudaError_t of the launch).
cudaError_t sas_launch_link_poses(const double *body, const SasPoseModel &m, int E, float *out, cudaStream_t st);
cudaError_t sas_launch_unproject(const float *depth, int I, int W, int H, float fx, float fy, float cx, float cy, float max_depth, float *xyz,
                                 unsigned char *valid, cudaStream_t st);
cudaError_t sas_launch_work_lists(const unsigned int *row_tile_count, const int *view_of_image, const unsigned int *tile_dyn_mask, int C, int I,
                                  int W, int H, int tile_w, int tile_h, SasWork *work, unsigned char *view_tile_needed, int *sort_list,
                                  int *comp_list, cudaStream_t st, int *n_launches);
// one-launch replacement of the two sas_launch_scan calls + sas_launch_work_lists for small problems
bool sas_scan_lists_fused_ok(int C, int I, int T);
cudaError_t sas_launch_scan_lists_fused(const unsigned int *row_tile_count, const int *view_of_image, const unsigned int *tile_dyn_mask, int C,
                                        int I, int W, int H, int tile_w, int tile_h, int *row_tile_offset, int *img_tile_offset,
                                        SasCounters *counters, long long cap_src, long long cap_all, SasWork *work,
                                        unsigned char *view_tile_needed, int *sort_list, int *comp_list, cudaStream_t st, int *n_launches);
// counters of this render -> ring[seq & 3], ++seq (device-side; the last node of a render, inside the CUDA graph)
struct SasSnapRing { SasCounters slot[4]; unsigned long long seq; };
cudaError_t sas_launch_snapshot(const SasCounters *counters, SasSnapRing *ring, cudaStream_t st);
cudaError_t sas_launch_project(const float4 *mean_op, const float4 *quat, const float4 *scale_link, const float *sh_planes,
                               int n_pad, int sh_degree, const float *link_poses, int L, const SasCamera *row_cams,
                               SasRowSpan span, int W, int H, int tile_w, int tile_h, SasRecord *rec, uint2 *bins,
                               uint2 *radii_out, unsigned int *row_tile_count, unsigned int *tile_dyn_mask, SasCounters *counters,
                               const float *extend /*per-Gaussian radius extend or null = 3.33*/, cudaStream_t st);
cudaError_t sas_bin_init();
cudaError_t sas_project_init();
cudaError_t sas_launch_scan(const unsigned int *counts, int n, const int *view_of_image, int C, int T, int *offsets,
                            unsigned int *block_sums, SasCounters *counters, long long capacity, cudaStream_t st, int *n_launches);
cudaError_t sas_launch_emit(const uint2 *bins, SasRowSpan span, int tile_w, int tile_h, const int *row_tile_offset,
                            unsigned int *row_tile_cursor, uint2 *pairs, long long capacity, cudaStream_t st);
cudaError_t sas_sort_init(int n_sms);
cudaError_t sas_binsort_init();
// the four shared-memory size classes as bin sorts (sas_binsort.cu); class c on stream s_of[c]
cudaError_t sas_launch_bin_sort(uint2 *pairs, const int *row_tile_offset, const int *sort_list, SasWork *work, long long capacity, int n_sms,
                                cudaStream_t *s_of);
// aux / ev: SAS_SORT_CLASSES - 1 side streams and 2 * SAS_SORT_CLASSES events for the fork / join (may be null: serial)
cudaError_t sas_launch_tile_sort(uint2 *pairs, uint2 *pairs_alt, const int *row_tile_offset, const int *sort_list, SasWork *work,
                                 long long capacity, const int *canon_of /*internal -> canonical id of background Gaussians, or null*/,
                                 int n_bg, int2 *big_items /*[big_items_cap] scratch: (bucket, chunk) of buckets beyond shared memory*/,
                                 int big_items_cap, int big_on /*enqueue the multi-CTA path for giant buckets*/, cudaStream_t st,
                                 cudaStream_t *aux, cudaEvent_t *ev, int *n_launches);
// fill_static: also write the ids of tiles whose link bucket is empty (the compositor never reads those)
// bg_reach / tile_dyn_mask / comp_len (all or none): compact every image tile's list to the records its compositor needs
cudaError_t sas_launch_merge(const uint2 *pairs, const int *row_tile_offset, const int *img_tile_offset, const int *view_of_image,
                             int C, int I, int T, int *merged_gid, long long cap_src, long long cap_all, int fill_static,
                             const int *comp_list, const SasWork *work, const unsigned *bg_reach, const unsigned *tile_dyn_mask,
                             int *comp_len, int vrec_codes, const SasRecord *rec, int *merged2 /*part-1 lists or null*/, int n_bg, int n_dyn,
                             int tile_w, int W, int H, cudaStream_t st);
cudaError_t sas_launch_bg_reach(const uint2 *pairs, const int *row_tile_offset, const SasRecord *rec, int n_bg, int C, int T, int tile_w,
                                long long cap_src, unsigned *bg_reach, SasRecord *vrec, int n_sms, cudaStream_t st);
cudaError_t sas_composite_init(int n_sms);
// Composites the ordered work list (image tiles with link Gaussians + the needed view tiles), then
// copies the view tiles into the images whose link bucket for that tile is empty.
cudaError_t sas_launch_composite(const SasRecord *rec, const int *merged_gid, const uint2 *pairs, const int *img_tile_offset,
                                 const int *row_tile_offset, const unsigned int *row_tile_count, const int *view_of_image,
                                 const unsigned int *tile_dyn_mask, const int *comp_len /*compacted list lengths or null*/, const SasRecord *vrec /*view bucket records or null*/, const int *merged2 /*part-1 lists or null*/,
                                 const int *comp_list, SasWork *work, int n_bg, int n_dyn, int C,
                                 int I, int W, int H, int tile_w, int tile_h, float bg0, float bg1, float bg2, SasOutputs img_out,
                                 SasViewBuf view_out, long long cap_src, long long cap_all, SasCounters *counters, cudaStream_t st,
                                 int *n_launches);
cudaError_t sas_launch_ns_post(float *rgb, float *depth, const float *alpha, uint8_t *rgb_u8, uint16_t *depth_f16, int I, int W, int H,
                               float bg0, float bg1, float bg2, unsigned int *img_max, cudaStream_t st, int *n_launches);
