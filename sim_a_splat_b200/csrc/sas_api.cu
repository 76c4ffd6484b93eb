// sas_api.cu -- C ABI of libsas_b200.so (see include/sas.h for the contract and the reference
// interfaces each entry point replaces).  Host-side only: buffer management, camera packing and
// view de-duplication, kernel sequencing on the caller's stream.  No CPU fallback exists:
// without a usable CUDA device every entry point fails with SAS_ECUDA.
#include "../../include/sas.h"
#include "sas_common.cuh"

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: stage ranges for nsys / ncu --nvtx filtering

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

}  // namespace

struct sas_renderer {
    int device = 0;
    int N = 0, n_bg = 0, n_dyn = 0, n_pad = 0, sh_degree = 0, L = 0;
    sas_config cfg{};
    std::string err;
    // scene (SoA)
    float4 *mean_op = nullptr, *quat = nullptr, *scale_link = nullptr;
    float *sh_planes = nullptr;
    float *extend = nullptr;             // per-Gaussian radius extend (radius_rule = cuda_1_5), else null
    // INTERNAL order of the background Gaussians: Morton order of their means (spatially close Gaussians sit next to
    // each other, so a CTA's 256 Gaussians project into neighbouring tiles and the emission writes each bucket in
    // contiguous runs).  canon_of[internal] = canonical index; null = the internal order is the canonical one.
    int *canon_of = nullptr;
    std::vector<int> h_canon;
    // per-call state
    float *link_poses = nullptr;
    int n_envs = 0;
    SasCamera *row_cams = nullptr;       // [C views | I images]
    int *view_of_image = nullptr;        // [I]
    unsigned char *stage_host = nullptr; // pinned staging: row cams + view_of_image
    size_t stage_bytes = 0;
    cudaEvent_t cams_copied = nullptr;
    std::vector<int> h_view_of_image;
    int I = 0, C = 0, W = 0, H = 0, tile_w = 0, tile_h = 0;
    bool cameras_set = false;
    // grow-only work buffers
    DevBuf rec, bins, row_tile_count, row_tile_cursor, row_tile_offset, img_tile_offset, block_sums, pairs, pairs_alt, merged, img_max;
    DevBuf sort_list, comp_list, view_needed, v_acc, v_depth, v_nc, radii, dyn_mask, big_items, bg_reach, comp_len, vrec, merged2;
    SasPoseModel pose_model{};            // device pointers inside; pose_model.L == 0: not set
    double *fk_dev = nullptr;             // [L,15]: fk_R, fk_t, body_offset
    DevBuf body_dev, unproj_d, unproj_xyz, unproj_valid;
    SasFkModel fk_model{};                // device pointers inside; fk_model.J == 0: not set
    int fk_actuated = 0;
    DevBuf fk_blob, fk_q, fk_extra;
    cudaStream_t aux[SAS_SORT_CLASSES - 1] = {};      // side streams: the bucket-sort size classes run concurrently
    cudaEvent_t fork_ev[SAS_SORT_CLASSES] = {};       // [0] fork, [1..] joins
    cudaEvent_t scan_ev[3] = {};                      // fork / two joins of the scan + work-list chains
    cudaEvent_t span_ev[2] = {};                      // fork / join of the background span next to the link span
    // device staging when the caller's outputs are host memory; set [1] only for SAS_MEM_HOST_ASYNC, which alternates
    // between the two so that the D2H copy of one render overlaps the next render
    DevBuf o_pack[2];                  // ONE block per set holding every requested host output (sas_packed_layout)
    DevBuf o_rgb[1], o_alpha[1];       // nerfstudio post-processing scratch when rgb / alpha are not requested themselves
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t render_done[2] = {}, copy_done[2] = {};
    unsigned long long async_seq = 0;    // async renders issued so far; render k uses staging set k & 1
    unsigned long long async_render_seq[2] = {0, 0};  // position of the async render of each set among all renders
    long long max_isects = 0;          // capacity of the merged id list (entries of 4 B)
    long long max_isects_src = 0;      // capacity of the bucket entries (pairs and their sort buffer, 8 B each)
    SasCounters *counters = nullptr;
    // overflow reporting without a sync per frame: every render snapshots its counters into a device ring; host-output
    // renders bring their slot back with the observation (sas_common.cuh SasSnapRing)
    SasSnapRing *snap = nullptr;
    SasSnapRing *h_snap = nullptr;       // pinned host copy
    unsigned long long render_seq = 0;   // renders issued (== ring->seq once they have run)
    unsigned long long ovf_floor = 0;    // renders before this one are not reported any more (reserve / poll)
    int max_env_index = 0;               // largest env an image of the current camera set refers to
    // Has this renderer seen tile lists long enough for the multi-CTA big-bucket sort (sas_sort.cu)?  Set from the
    // counters the host reads anyway (stats, overflow polls, snapshots that travel with host outputs); until then the
    // step's graph does not carry that path's launches.  A hint: correctness does not depend on it.
    bool big_hint = false;
    bool rendered = false;
    int last_launches = 0;
    bool stage_timing = false;
    cudaEvent_t stage_ev[SAS_N_STAGES + 1] = {};
    // CUDA graph of the device work of one sas_render (captured on first use, replayed while the key matches)
    struct GraphKey {
        int I, C, W, H, ns, flags, big;
        float bg[3];
        void *rgb, *depth, *alpha, *u8, *nc, *f16;
        long long max_isects, max_isects_src;
        unsigned long long buf_gen;
    };
    // two cached graphs (the async host path alternates between two output sets), least recently used replaced
    struct GraphSlot {
        cudaGraphExec_t exec = nullptr;
        GraphKey key{};
        int launches = 0;
        unsigned long long used = 0;
    } graphs[2];
    GraphKey pending_key[2]{};       // a graph is only built for a key seen before (one of the last two unseen keys)
    bool pending_valid[2] = {false, false};
    int pending_next = 0;
    unsigned long long graph_clock = 0;
    unsigned long long buf_gen = 0;  // bumped whenever a work buffer is (re)allocated: invalidates the graph
    int graph_mode = -1;             // SAS_GRAPH environment variable (default on)
};

namespace {

int fail(sas_renderer *r, int code, const std::string &msg) {
    if (r) r->err = msg; else g_create_error = msg;
    return code;
}

#define SAS_CUDA(r, call)                                                                                  \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            char buf_[512];                                                                                \
            snprintf(buf_, sizeof(buf_), "%.300s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return fail((r), e_ == cudaErrorMemoryAllocation ? SAS_ENOMEM : SAS_ECUDA, buf_);              \
        }                                                                                                  \
    } while (0)

int ensure(sas_renderer *r, DevBuf &b, size_t bytes) {
    if (b.bytes >= bytes && b.p) return SAS_OK;
    if (b.p) {
        SAS_CUDA(r, cudaDeviceSynchronize());
        SAS_CUDA(r, cudaFree(b.p));
        b.p = nullptr; b.bytes = 0;
    }
    if (bytes == 0) bytes = 16;
    SAS_CUDA(r, cudaMalloc(&b.p, bytes));
    b.bytes = bytes;
    r->buf_gen++;
    return SAS_OK;
}

void free_buf(DevBuf &b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.bytes = 0;
}

// m_src bucket entries (pairs + sort buffer), m_all merged ids; 0 = leave as it is.  Offsets are int32.
int ensure_isects(sas_renderer *r, long long m_src, long long m_all) {
    if (m_src > 0x7ffffff0ll) m_src = 0x7ffffff0ll;
    if (m_all > 0x7ffffff0ll) m_all = 0x7ffffff0ll;
    int rc;
    if (m_src > r->max_isects_src || (m_src > 0 && !r->pairs.p)) {
        if ((rc = ensure(r, r->pairs, (size_t)m_src * sizeof(uint2)))) return rc;
        if ((rc = ensure(r, r->pairs_alt, (size_t)m_src * sizeof(uint2)))) return rc;
        r->max_isects_src = m_src;
    }
    if (m_all > r->max_isects || (m_all > 0 && !r->merged.p)) {
        if ((rc = ensure(r, r->merged, (size_t)m_all * sizeof(int)))) return rc;
        r->max_isects = m_all;
    }
    r->ovf_floor = r->render_seq;  // earlier overflows have been answered
    return SAS_OK;
}

struct PackedLayout {
    size_t off[6] = {0, 0, 0, 0, 0, 0};
    size_t total = 0;
};
// order and element sizes of the SAS_OUT_* bits: u8 rgb (3), f16 depth (2), f32 depth (4), alpha (4), rgb (12), n_contrib (4)
PackedLayout packed_layout(size_t P, int want) {
    static const size_t kBytes[6] = {3, 2, 4, 4, 12, 4};
    PackedLayout l;
    for (int k = 0; k < 6; k++)
        if (want & (1 << k)) {
            l.off[k] = l.total;
            l.total = (l.total + P * kBytes[k] + 255) & ~(size_t)255;
        }
    return l;
}

void note_counters(sas_renderer *r, const SasCounters &c) {
    if ((long long)c.max_tile_list > SAS_SORT_CAP3) r->big_hint = true;
}

int overflow_error(sas_renderer *r, const SasCounters &c, const char *who) {
    char buf[320];
    snprintf(buf, sizeof(buf), "%s: the frame is INCOMPLETE, intersection buffers overflowed (bucket entries %llu of capacity %lld, merged "
             "ids %llu of capacity %lld): sas_reserve_isects2 and render again", who, c.n_isects_src, r->max_isects_src, c.n_isects, r->max_isects);
    return fail(r, SAS_ECAPACITY, buf);
}

void fill_camera(SasCamera &c, const float *vm, const float *K, int width, int height, int env) {
    for (int k = 0; k < 12; k++) c.vm[k] = vm[k];
    c.fx = K[0]; c.fy = K[4]; c.cx = K[2]; c.cy = K[5];
    // camera origin -R^T t in fp64, rounded once (the oracle does the same on the same fp32 viewmat)
    for (int a = 0; a < 3; a++) {
        double s = 0.0;
        for (int b = 0; b < 3; b++) s += (double)vm[4 * b + a] * (double)vm[4 * b + 3];
        c.campos[a] = (float)(-s);
    }
    c.env = env;
    // same fp32 operation sequence as oracle/splat_oracle.c:project_one (host code is built with
    // -ffp-contract=off, SSE fp32: every operation individually rounded)
    const volatile float Wf = (float)width, Hf = (float)height;
    const volatile float tan_fovx = 0.5f * Wf / c.fx, tan_fovy = 0.5f * Hf / c.fy;
    const volatile float ax = 0.3f * tan_fovx, ay = 0.3f * tan_fovy;
    c.lim_x_pos = (Wf - c.cx) / c.fx + ax;
    c.lim_x_neg = c.cx / c.fx + ax;
    c.lim_y_pos = (Hf - c.cy) / c.fy + ay;
    c.lim_y_neg = c.cy / c.fy + ay;
}

}  // namespace

// compacted compositor lists are used when a view is shared by at least this many images (SAS_COMPACT_RATIO)
static int compact_min_ratio() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SAS_COMPACT_RATIO");
        v = e ? atoi(e) : 4;
        if (v < 1) v = 1;
    }
    return v;
}

int sas_variant_flags() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SAS_VARIANT");
        v = e ? (int)strtol(e, nullptr, 0) & 0x7fff : 0x32cf;
    }
    return v;
}

extern "C" {

const char *sas_version(void) { return "sas_b200 0.4.0 (sm_100a)"; }

int sas_packed_layout(int n_images, int width, int height, int want, size_t offsets[6], size_t *total_bytes) {
    if (n_images <= 0 || width <= 0 || height <= 0 || !offsets || !total_bytes || (want & ~63)) return SAS_EINVAL;
    const PackedLayout l = packed_layout((size_t)n_images * width * height, want);
    for (int k = 0; k < 6; k++)
        if (want & (1 << k)) offsets[k] = l.off[k];
    *total_bytes = l.total;
    return SAS_OK;
}

const char *sas_last_error(const sas_renderer *r) { return r ? r->err.c_str() : g_create_error.c_str(); }

int sas_last_launch_count(const sas_renderer *r) { return r ? r->last_launches : 0; }

int sas_create(const sas_scene_desc *scene, const sas_config *cfg, int device, sas_renderer **out) {
    if (!scene || !cfg || !out) return fail(nullptr, SAS_EINVAL, "sas_create: null argument");
    *out = nullptr;
    if (scene->n_gaussians <= 0 || scene->sh_degree < 0 || scene->sh_degree > 3 || scene->n_links < 0 || !scene->means ||
        !scene->quats || !scene->scales || !scene->opacities || !scene->sh)
        return fail(nullptr, SAS_EINVAL, "sas_create: bad scene descriptor");
    if (cfg->max_width > 4080 || cfg->max_height > 4080 || cfg->max_images <= 0 || cfg->max_images > 32767)
        return fail(nullptr, SAS_EINVAL, "sas_create: max_width/max_height must be <= 4080 and 0 < max_images <= 32767");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        char buf[256];
        snprintf(buf, sizeof(buf), "sas_create: no usable CUDA device %d (%s); there is no CPU fallback", device,
                 e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range");
        return fail(nullptr, SAS_ECUDA, buf);
    }
    sas_renderer *r = new sas_renderer();
    r->device = device;
    r->cfg = *cfg;
    const int N = scene->n_gaussians, K = (scene->sh_degree + 1) * (scene->sh_degree + 1);
    r->N = N; r->sh_degree = scene->sh_degree; r->L = scene->n_links;
    r->n_pad = (N + 31) & ~31;
    auto bail = [&](int code) { std::string m = r->err; sas_destroy(r); g_create_error = m; return code; };
#define SAS_TRY(x) do { int rc_ = (x); if (rc_) return bail(rc_); } while (0)
#define SAS_CUDA_C(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { r->err = std::string(#call) + " failed: " + cudaGetErrorString(e_); return bail(e_ == cudaErrorMemoryAllocation ? SAS_ENOMEM : SAS_ECUDA); } } while (0)
    SAS_CUDA_C(cudaSetDevice(device));
    int n_sms = 0;
    SAS_CUDA_C(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device));
    SAS_CUDA_C(sas_sort_init(n_sms));
    SAS_CUDA_C(sas_composite_init(n_sms));
    SAS_CUDA_C(sas_bin_init());
    SAS_CUDA_C(sas_project_init());

    // bring the scene to the host if needed, pack SoA, upload
    std::vector<float> h_means((size_t)N * 3), h_quats((size_t)N * 4), h_scales((size_t)N * 3), h_op(N), h_sh((size_t)N * K * 3);
    std::vector<int32_t> h_link(N, -1);
    const cudaMemcpyKind kind = scene->mem == SAS_MEM_DEVICE ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost;
    SAS_CUDA_C(cudaMemcpy(h_means.data(), scene->means, h_means.size() * 4, kind));
    SAS_CUDA_C(cudaMemcpy(h_quats.data(), scene->quats, h_quats.size() * 4, kind));
    SAS_CUDA_C(cudaMemcpy(h_scales.data(), scene->scales, h_scales.size() * 4, kind));
    SAS_CUDA_C(cudaMemcpy(h_op.data(), scene->opacities, h_op.size() * 4, kind));
    SAS_CUDA_C(cudaMemcpy(h_sh.data(), scene->sh, h_sh.size() * 4, kind));
    if (scene->link_id && scene->n_links > 0) SAS_CUDA_C(cudaMemcpy(h_link.data(), scene->link_id, (size_t)N * 4, kind));
    // canonical order: all background Gaussians (link_id < 0) first, then the link copies
    int n_bg = 0;
    while (n_bg < N && h_link[n_bg] < 0) n_bg++;
    for (int g = n_bg; g < N; g++)
        if (h_link[g] < 0 || h_link[g] >= scene->n_links) {
            r->err = "sas_create: scene is not in canonical order (background first, then links) or link_id out of range";
            return bail(SAS_EINVAL);
        }
    r->n_bg = n_bg; r->n_dyn = N - n_bg;

    // Optional internal order (SAS_MORTON=1): background by the Morton code of the mean (10 bits per axis over the
    // background's bounding box, canonical index as the tie-break), link Gaussians as they are.  Measured (round 2b,
    // profiles/r02b_morton_order.txt): emission of the 6 M-Gaussian 1080p chunk 11.8 -> 9.8 ms, but the Gaussians of a CTA
    // then hit the SAME few tiles and serialise on the shared-memory histogram atomics of projection and emission --
    // configs[2] emission 0.065 -> 0.192 ms, projection 0.080 -> 0.110 ms.  Default: only for large scenes (below).
    std::vector<int> order(N);
    for (int g = 0; g < N; g++) order[g] = g;
    {
        const char *em = getenv("SAS_MORTON");
        const char *es = getenv("SAS_SORT");
        // unset: on for scenes of 2 M background Gaussians and more (the regime where it was measured to pay)
        const bool on = em ? atoi(em) != 0 : n_bg >= 2000000;
        const bool want = on && !(es && es[0] == 'b') && n_bg > 1;
        if (want) {
            float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
            for (int g = 0; g < n_bg; g++)
                for (int a = 0; a < 3; a++) {
                    const float v = h_means[3 * g + a];
                    if (v == v) { lo[a] = v < lo[a] ? v : lo[a]; hi[a] = v > hi[a] ? v : hi[a]; }
                }
            std::vector<unsigned long long> code(n_bg);
            auto spread = [](unsigned v) {  // 10 bits -> every third bit
                unsigned long long x = v & 0x3ffu;
                x = (x | (x << 16)) & 0x030000ffull;
                x = (x | (x << 8)) & 0x0300f00full;
                x = (x | (x << 4)) & 0x030c30c3ull;
                x = (x | (x << 2)) & 0x09249249ull;
                return x;
            };
            for (int g = 0; g < n_bg; g++) {
                unsigned q[3];
                for (int a = 0; a < 3; a++) {
                    const float ext = hi[a] - lo[a];
                    float t = ext > 0.0f ? (h_means[3 * g + a] - lo[a]) / ext : 0.0f;
                    if (!(t >= 0.0f)) t = 0.0f;
                    if (t > 1.0f) t = 1.0f;
                    q[a] = (unsigned)(t * 1023.0f);
                }
                code[g] = ((spread(q[0]) | (spread(q[1]) << 1) | (spread(q[2]) << 2)) << 32) | (unsigned)g;
            }
            std::sort(code.begin(), code.end());
            r->h_canon.resize(n_bg);
            for (int i = 0; i < n_bg; i++) { order[i] = (int)(code[i] & 0xffffffffull); r->h_canon[i] = order[i]; }
        }
    }
    std::vector<float4> p_mean(N), p_quat(N), p_scale(N);
    std::vector<float> p_sh((size_t)K * 3 * r->n_pad, 0.0f);
    for (int i = 0; i < N; i++) {
        const int g = order[i];  // canonical index of the Gaussian at internal position i
        p_mean[i] = make_float4(h_means[3 * g], h_means[3 * g + 1], h_means[3 * g + 2], h_op[g]);
        p_quat[i] = make_float4(h_quats[4 * g], h_quats[4 * g + 1], h_quats[4 * g + 2], h_quats[4 * g + 3]);
        float lid;
        int32_t li = h_link[g] < 0 ? -1 : h_link[g];
        memcpy(&lid, &li, 4);
        p_scale[i] = make_float4(h_scales[3 * g], h_scales[3 * g + 1], h_scales[3 * g + 2], lid);
        for (int k = 0; k < K; k++)
            for (int ch = 0; ch < 3; ch++) p_sh[(size_t)(3 * k + ch) * r->n_pad + i] = h_sh[((size_t)g * K + k) * 3 + ch];
    }
    if (!r->h_canon.empty()) {
        SAS_CUDA_C(cudaMalloc(&r->canon_of, sizeof(int) * n_bg));
        SAS_CUDA_C(cudaMemcpy(r->canon_of, r->h_canon.data(), sizeof(int) * n_bg, cudaMemcpyHostToDevice));
    }
    SAS_CUDA_C(cudaMalloc(&r->mean_op, sizeof(float4) * N));
    SAS_CUDA_C(cudaMalloc(&r->quat, sizeof(float4) * N));
    SAS_CUDA_C(cudaMalloc(&r->scale_link, sizeof(float4) * N));
    SAS_CUDA_C(cudaMalloc(&r->sh_planes, sizeof(float) * p_sh.size()));
    SAS_CUDA_C(cudaMemcpy(r->mean_op, p_mean.data(), sizeof(float4) * N, cudaMemcpyHostToDevice));
    SAS_CUDA_C(cudaMemcpy(r->quat, p_quat.data(), sizeof(float4) * N, cudaMemcpyHostToDevice));
    SAS_CUDA_C(cudaMemcpy(r->scale_link, p_scale.data(), sizeof(float4) * N, cudaMemcpyHostToDevice));
    SAS_CUDA_C(cudaMemcpy(r->sh_planes, p_sh.data(), sizeof(float) * p_sh.size(), cudaMemcpyHostToDevice));
    if (cfg->radius_rule == SAS_RADIUS_CUDA_1_5) {
        // the same double-precision expression, rounded once to float, as oracle/splat_oracle.c:orc_extend
        std::vector<float> ext(N);
        const float amin = 1.0f / 255.0f;
        for (int i = 0; i < N; i++) {
            const float op = h_op[order[i]];
            float e = 0.0f;
            if (!(op < amin)) {
                e = (float)sqrt(2.0 * log((double)op / (double)amin));
                if (!(e < 3.33f)) e = 3.33f;
            }
            ext[i] = e;
        }
        SAS_CUDA_C(cudaMalloc(&r->extend, sizeof(float) * N));
        SAS_CUDA_C(cudaMemcpy(r->extend, ext.data(), sizeof(float) * N, cudaMemcpyHostToDevice));
    } else if (cfg->radius_rule != SAS_RADIUS_TORCH) {
        r->err = "sas_create: unknown radius_rule";
        return bail(SAS_EINVAL);
    }

    const int maxE = cfg->max_envs > 0 ? cfg->max_envs : 1;
    if (r->L > 0) {
        // identity until sas_set_link_poses: an env that has never been given poses renders its links at rest
        std::vector<float> ident((size_t)12 * maxE * r->L, 0.0f);
        for (size_t k = 0; k < (size_t)maxE * r->L; k++) ident[12 * k] = ident[12 * k + 4] = ident[12 * k + 8] = 1.0f;
        SAS_CUDA_C(cudaMalloc(&r->link_poses, sizeof(float) * ident.size()));
        SAS_CUDA_C(cudaMemcpy(r->link_poses, ident.data(), sizeof(float) * ident.size(), cudaMemcpyHostToDevice));
    }
    SAS_CUDA_C(cudaMalloc(&r->row_cams, sizeof(SasCamera) * 2 * cfg->max_images));
    SAS_CUDA_C(cudaMalloc(&r->view_of_image, sizeof(int) * cfg->max_images));
    r->stage_bytes = sizeof(SasCamera) * 2 * cfg->max_images + sizeof(int) * cfg->max_images;
    SAS_CUDA_C(cudaMallocHost(&r->stage_host, r->stage_bytes));
    SAS_CUDA_C(cudaEventCreateWithFlags(&r->cams_copied, cudaEventDisableTiming));
    for (cudaStream_t &s : r->aux) SAS_CUDA_C(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    for (cudaEvent_t &ev : r->fork_ev) SAS_CUDA_C(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (cudaEvent_t &ev : r->scan_ev) SAS_CUDA_C(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (cudaEvent_t &ev : r->span_ev) SAS_CUDA_C(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    SAS_CUDA_C(cudaMalloc(&r->counters, sizeof(SasCounters) + sizeof(SasWork)));  // SasWork follows the counters
    SAS_CUDA_C(cudaMemset(r->counters, 0, sizeof(SasCounters) + sizeof(SasWork)));
    SAS_CUDA_C(cudaMalloc(&r->snap, sizeof(SasSnapRing)));
    SAS_CUDA_C(cudaMemset(r->snap, 0, sizeof(SasSnapRing)));
    SAS_CUDA_C(cudaMallocHost(&r->h_snap, sizeof(SasSnapRing)));
    memset(r->h_snap, 0, sizeof(SasSnapRing));
    const long long m = cfg->max_isects > 0 ? cfg->max_isects : (long long)1 << 22;
    SAS_TRY(ensure_isects(r, cfg->max_isects_src > 0 ? cfg->max_isects_src : m, m));
#undef SAS_TRY
#undef SAS_CUDA_C
    *out = r;
    return SAS_OK;
}

void sas_destroy(sas_renderer *r) {
    if (!r) return;
    cudaSetDevice(r->device);
    cudaDeviceSynchronize();
    cudaFree(r->mean_op); cudaFree(r->quat); cudaFree(r->scale_link); cudaFree(r->sh_planes);
    cudaFree(r->fk_dev); cudaFree(r->extend); cudaFree(r->canon_of);
    cudaFree(r->link_poses); cudaFree(r->row_cams); cudaFree(r->view_of_image); cudaFree(r->counters); cudaFree(r->snap);
    if (r->h_snap) cudaFreeHost(r->h_snap);
    if (r->stage_host) cudaFreeHost(r->stage_host);
    if (r->cams_copied) cudaEventDestroy(r->cams_copied);
    for (cudaEvent_t ev : r->stage_ev)
        if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : r->fork_ev)
        if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : r->scan_ev)
        if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : r->span_ev)
        if (ev) cudaEventDestroy(ev);
    for (auto &g : r->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    for (cudaStream_t s : r->aux)
        if (s) cudaStreamDestroy(s);
    if (r->copy_stream) cudaStreamDestroy(r->copy_stream);
    for (cudaEvent_t ev : r->render_done)
        if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : r->copy_done)
        if (ev) cudaEventDestroy(ev);
    DevBuf *bufs[] = {&r->rec, &r->bins, &r->row_tile_count, &r->row_tile_cursor, &r->row_tile_offset, &r->img_tile_offset, &r->block_sums,
                      &r->pairs, &r->pairs_alt, &r->merged, &r->img_max, &r->o_rgb[0], &r->o_alpha[0], &r->o_pack[0], &r->o_pack[1],
                      &r->big_items, &r->bg_reach, &r->comp_len, &r->vrec, &r->merged2, &r->sort_list, &r->comp_list, &r->view_needed, &r->v_acc, &r->v_depth, &r->v_nc, &r->radii, &r->dyn_mask, &r->body_dev, &r->unproj_d, &r->unproj_xyz, &r->unproj_valid,
                      &r->fk_blob, &r->fk_q, &r->fk_extra};
    for (DevBuf *b : bufs) free_buf(*b);
    delete r;
}

int sas_reserve_isects(sas_renderer *r, long long max_isects) {
    if (!r || max_isects <= 0) return fail(r, SAS_EINVAL, "sas_reserve_isects: bad argument");
    SAS_CUDA(r, cudaSetDevice(r->device));
    return ensure_isects(r, max_isects, max_isects);
}

int sas_reserve_isects2(sas_renderer *r, long long max_src, long long max_all) {
    if (!r || max_src < 0 || max_all < 0) return fail(r, SAS_EINVAL, "sas_reserve_isects2: bad argument");
    SAS_CUDA(r, cudaSetDevice(r->device));
    return ensure_isects(r, max_src, max_all);
}

int sas_set_link_poses(sas_renderer *r, const float *poses, int n_envs, int mem, void *stream) {
    if (!r) return SAS_EINVAL;
    if (r->L == 0) return SAS_OK;
    if (!poses || n_envs <= 0) return fail(r, SAS_EINVAL, "sas_set_link_poses: null poses or n_envs <= 0");
    const int maxE = r->cfg.max_envs > 0 ? r->cfg.max_envs : 1;
    if (n_envs > maxE) return fail(r, SAS_ECAPACITY, "sas_set_link_poses: n_envs exceeds max_envs");
    SAS_CUDA(r, cudaSetDevice(r->device));
    SAS_CUDA(r, cudaMemcpyAsync(r->link_poses, poses, sizeof(float) * 12 * (size_t)n_envs * r->L,
                                mem == SAS_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, (cudaStream_t)stream));
    r->n_envs = n_envs;
    return SAS_OK;
}

int sas_set_pose_model(sas_renderer *r, const sas_pose_model *model) {
    if (!r || !model) return SAS_EINVAL;
    if (model->n_links != r->L || r->L <= 0 || !model->fk_R || !model->fk_t)
        return fail(r, SAS_EINVAL, "sas_set_pose_model: n_links must equal the scene's link count (> 0) and fk_R / fk_t must be given");
    SAS_CUDA(r, cudaSetDevice(r->device));
    if (!r->fk_dev) SAS_CUDA(r, cudaMalloc(&r->fk_dev, sizeof(double) * 15 * (size_t)r->L));
    SAS_CUDA(r, cudaMemcpy(r->fk_dev, model->fk_R, sizeof(double) * 9 * (size_t)r->L, cudaMemcpyHostToDevice));
    SAS_CUDA(r, cudaMemcpy(r->fk_dev + 9 * (size_t)r->L, model->fk_t, sizeof(double) * 3 * (size_t)r->L, cudaMemcpyHostToDevice));
    if (model->body_offset)
        SAS_CUDA(r, cudaMemcpy(r->fk_dev + 12 * (size_t)r->L, model->body_offset, sizeof(double) * 3 * (size_t)r->L, cudaMemcpyHostToDevice));
    else
        SAS_CUDA(r, cudaMemset(r->fk_dev + 12 * (size_t)r->L, 0, sizeof(double) * 3 * (size_t)r->L));
    SasPoseModel &m = r->pose_model;
    m.L = r->L;
    m.s = model->scale;
    for (int i = 0; i < 9; i++) m.Ri[i] = model->Ri[i];
    for (int i = 0; i < 3; i++) m.ti[i] = model->ti[i];
    m.fk_R = r->fk_dev;
    m.fk_t = r->fk_dev + 9 * (size_t)r->L;
    m.offset = r->fk_dev + 12 * (size_t)r->L;
    return SAS_OK;
}

int sas_set_link_poses_from_bodies(sas_renderer *r, const double *body, int n_envs, int mem, void *stream) {
    if (!r) return SAS_EINVAL;
    if (r->pose_model.L == 0) return fail(r, SAS_ESTATE, "sas_set_link_poses_from_bodies: sas_set_pose_model has not been called");
    if (!body || n_envs <= 0) return fail(r, SAS_EINVAL, "sas_set_link_poses_from_bodies: null poses or n_envs <= 0");
    const int maxE = r->cfg.max_envs > 0 ? r->cfg.max_envs : 1;
    if (n_envs > maxE) return fail(r, SAS_ECAPACITY, "sas_set_link_poses_from_bodies: n_envs exceeds max_envs");
    SAS_CUDA(r, cudaSetDevice(r->device));
    cudaStream_t st = (cudaStream_t)stream;
    const double *d_body = body;
    if (mem != SAS_MEM_DEVICE) {
        const size_t bytes = sizeof(double) * 7 * (size_t)n_envs * r->L;
        int rc;
        if ((rc = ensure(r, r->body_dev, bytes))) return rc;
        SAS_CUDA(r, cudaMemcpyAsync(r->body_dev.p, body, bytes, cudaMemcpyHostToDevice, st));
        d_body = (const double *)r->body_dev.p;
    }
    SAS_CUDA(r, sas_launch_link_poses(d_body, r->pose_model, n_envs, r->link_poses, st));
    r->n_envs = n_envs;
    return SAS_OK;
}

int sas_set_fk_model(sas_renderer *r, const sas_fk_model *model) {
    if (!r || !model) return SAS_EINVAL;
    const int J = model->n_joints, B = model->n_bodies;
    if (J <= 0 || J > sas_fk_max_joints() || B <= 0 || B > r->L || model->n_actuated <= 0 || !model->parent || !model->type || !model->qindex ||
        !model->mimic || !model->origin || !model->axis || !model->body_frame)
        return fail(r, SAS_EINVAL, "sas_set_fk_model: 0 < n_joints <= 48, 0 < n_bodies <= the scene's link count, and all arrays must be given");
    for (int j = 0; j < J; j++)
        if (model->parent[j] < 0 || model->parent[j] > j || model->qindex[j] >= model->n_actuated || model->type[j] < 0 || model->type[j] > 2)
            return fail(r, SAS_EINVAL, "sas_set_fk_model: joints must be in parent-before-child order with valid type / qindex");
    for (int b = 0; b < B; b++)
        if (model->body_frame[b] < 0 || model->body_frame[b] > J) return fail(r, SAS_EINVAL, "sas_set_fk_model: body_frame out of range");
    SAS_CUDA(r, cudaSetDevice(r->device));
    // one blob: ints (parent, type, qindex | body_frame) then doubles (mimic, origin 12 per joint, axis)
    const size_t n_int = (size_t)3 * J + B, int_bytes = (n_int * 4 + 7) & ~(size_t)7, n_dbl = (size_t)J * (2 + 12 + 3);
    std::vector<unsigned char> blob(int_bytes + n_dbl * 8);
    int *pi = reinterpret_cast<int *>(blob.data());
    double *pd = reinterpret_cast<double *>(blob.data() + int_bytes);
    for (int j = 0; j < J; j++) { pi[j] = model->parent[j]; pi[J + j] = model->type[j]; pi[2 * J + j] = model->qindex[j]; }
    for (int b = 0; b < B; b++) pi[3 * J + b] = model->body_frame[b];
    for (int j = 0; j < J; j++) {
        pd[2 * j] = model->mimic[2 * j]; pd[2 * j + 1] = model->mimic[2 * j + 1];
        double *o = pd + 2 * J + 12 * j;
        const double *M = model->origin + 16 * j;
        for (int i = 0; i < 3; i++) { for (int k = 0; k < 3; k++) o[3 * i + k] = M[4 * i + k]; o[9 + i] = M[4 * i + 3]; }
        for (int k = 0; k < 3; k++) pd[2 * J + 12 * J + 3 * j + k] = model->axis[3 * j + k];
    }
    int rc;
    if ((rc = ensure(r, r->fk_blob, blob.size()))) return rc;
    SAS_CUDA(r, cudaMemcpy(r->fk_blob.p, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    SasFkModel &m = r->fk_model;
    const unsigned char *d = (const unsigned char *)r->fk_blob.p;
    m.J = J; m.B = B;
    m.parent = (const int *)d; m.type = m.parent + J; m.qindex = m.parent + 2 * J; m.body_frame = m.parent + 3 * J;
    m.mimic = (const double *)(d + int_bytes); m.origin = m.mimic + 2 * J; m.axis = m.origin + 12 * J;
    for (int i = 0; i < 3; i++) {
        for (int k = 0; k < 3; k++) m.base[3 * i + k] = model->base ? model->base[4 * i + k] : (i == k ? 1.0 : 0.0);
        m.base[9 + i] = model->base ? model->base[4 * i + 3] : 0.0;
    }
    r->fk_actuated = model->n_actuated;
    return SAS_OK;
}

int sas_set_link_poses_from_joints(sas_renderer *r, const double *q, int n_envs, const double *free_bodies, int mem, void *stream) {
    if (!r) return SAS_EINVAL;
    if (r->fk_model.J == 0 || r->pose_model.L == 0)
        return fail(r, SAS_ESTATE, "sas_set_link_poses_from_joints: sas_set_fk_model and sas_set_pose_model must be called first");
    if (!q || n_envs <= 0) return fail(r, SAS_EINVAL, "sas_set_link_poses_from_joints: null joint angles or n_envs <= 0");
    const int maxE = r->cfg.max_envs > 0 ? r->cfg.max_envs : 1;
    if (n_envs > maxE) return fail(r, SAS_ECAPACITY, "sas_set_link_poses_from_joints: n_envs exceeds max_envs");
    SAS_CUDA(r, cudaSetDevice(r->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    const double *d_q = q, *d_extra = free_bodies;
    const int n_free = r->L - r->fk_model.B;
    if (mem != SAS_MEM_DEVICE) {
        const size_t qb = sizeof(double) * (size_t)n_envs * r->fk_actuated;
        if ((rc = ensure(r, r->fk_q, qb))) return rc;
        SAS_CUDA(r, cudaMemcpyAsync(r->fk_q.p, q, qb, cudaMemcpyHostToDevice, st));
        d_q = (const double *)r->fk_q.p;
        if (free_bodies && n_free > 0) {
            const size_t eb = sizeof(double) * 7 * (size_t)n_envs * n_free;
            if ((rc = ensure(r, r->fk_extra, eb))) return rc;
            SAS_CUDA(r, cudaMemcpyAsync(r->fk_extra.p, free_bodies, eb, cudaMemcpyHostToDevice, st));
            d_extra = (const double *)r->fk_extra.p;
        }
    }
    if ((rc = ensure(r, r->body_dev, sizeof(double) * 7 * (size_t)n_envs * r->L))) return rc;
    SAS_CUDA(r, sas_launch_fk(d_q, r->fk_actuated, r->fk_model, n_envs, n_free > 0 ? d_extra : nullptr, r->L, (double *)r->body_dev.p, st));
    SAS_CUDA(r, sas_launch_link_poses((const double *)r->body_dev.p, r->pose_model, n_envs, r->link_poses, st));
    r->n_envs = n_envs;
    return SAS_OK;
}

int sas_unproject_depth(sas_renderer *r, const float *depth, int n_images, int width, int height, const float *K, float max_depth,
                        float *xyz, uint8_t *valid, int mem, void *stream) {
    if (!r) return SAS_EINVAL;
    if (!depth || !xyz || !K || n_images <= 0 || width <= 0 || height <= 0) return fail(r, SAS_EINVAL, "sas_unproject_depth: bad argument");
    SAS_CUDA(r, cudaSetDevice(r->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t P = (size_t)n_images * width * height;
    const float *d_depth = depth;
    float *d_xyz = xyz;
    uint8_t *d_valid = valid;
    if (mem != SAS_MEM_DEVICE) {
        int rc;
        if ((rc = ensure(r, r->unproj_d, P * 4))) return rc;
        if ((rc = ensure(r, r->unproj_xyz, P * 12))) return rc;
        if (valid && (rc = ensure(r, r->unproj_valid, P))) return rc;
        SAS_CUDA(r, cudaMemcpyAsync(r->unproj_d.p, depth, P * 4, cudaMemcpyHostToDevice, st));
        d_depth = (const float *)r->unproj_d.p; d_xyz = (float *)r->unproj_xyz.p; d_valid = valid ? (uint8_t *)r->unproj_valid.p : nullptr;
    }
    SAS_CUDA(r, sas_launch_unproject(d_depth, n_images, width, height, K[0], K[4], K[2], K[5], max_depth, d_xyz, d_valid, st));
    if (mem != SAS_MEM_DEVICE) {
        SAS_CUDA(r, cudaMemcpyAsync(xyz, d_xyz, P * 12, cudaMemcpyDeviceToHost, st));
        if (valid) SAS_CUDA(r, cudaMemcpyAsync(valid, d_valid, P, cudaMemcpyDeviceToHost, st));
        SAS_CUDA(r, cudaStreamSynchronize(st));
    }
    return SAS_OK;
}

int sas_set_cameras(sas_renderer *r, const float *viewmats, const float *Ks, const int32_t *env_of_image, int n_images,
                    int width, int height, void *stream) {
    if (!r) return SAS_EINVAL;
    if (!viewmats || !Ks || n_images <= 0 || width <= 0 || height <= 0) return fail(r, SAS_EINVAL, "sas_set_cameras: bad argument");
    if (n_images > r->cfg.max_images) return fail(r, SAS_ECAPACITY, "sas_set_cameras: n_images exceeds max_images");
    if (width > r->cfg.max_width || height > r->cfg.max_height || width > 4080 || height > 4080)
        return fail(r, SAS_ECAPACITY, "sas_set_cameras: image size exceeds max_width/max_height");
    const int maxE = r->cfg.max_envs > 0 ? r->cfg.max_envs : 1;
    int max_env = 0;
    for (int i = 0; i < n_images; i++) {
        const int env = env_of_image ? env_of_image[i] : 0;
        if (env < 0 || (r->L > 0 && env >= maxE)) return fail(r, SAS_EINVAL, "sas_set_cameras: env_of_image out of range");
        if (env > max_env) max_env = env;
    }
    SAS_CUDA(r, cudaSetDevice(r->device));
    SAS_CUDA(r, cudaEventSynchronize(r->cams_copied));  // previous staging copy finished

    // Views: images with bit-identical (viewmat, K) share one view; the background is projected,
    // binned and sorted once per view.  Byte-wise comparison against the views found so far
    // (a handful in practice; the scan is bounded to keep this O(I)).
    std::vector<int> view_first;  // image index that defined each view
    r->h_view_of_image.assign(n_images, 0);
    for (int i = 0; i < n_images; i++) {
        int v = -1;
        const size_t nv = view_first.size();
        for (size_t k = nv > 64 ? nv - 64 : 0; k < nv; k++) {
            const int j = view_first[k];
            if (memcmp(viewmats + 16 * (size_t)i, viewmats + 16 * (size_t)j, 64) == 0 && memcmp(Ks + 9 * (size_t)i, Ks + 9 * (size_t)j, 36) == 0) {
                v = (int)k;
                break;
            }
        }
        if (v < 0) { v = (int)view_first.size(); view_first.push_back(i); }
        r->h_view_of_image[i] = v;
    }
    const int C = (int)view_first.size();
    SasCamera *cams = reinterpret_cast<SasCamera *>(r->stage_host);
    int *voi = reinterpret_cast<int *>(r->stage_host + sizeof(SasCamera) * (size_t)(C + n_images));
    for (int v = 0; v < C; v++) fill_camera(cams[v], viewmats + 16 * (size_t)view_first[v], Ks + 9 * (size_t)view_first[v], width, height, 0);
    for (int i = 0; i < n_images; i++) {
        fill_camera(cams[C + i], viewmats + 16 * (size_t)i, Ks + 9 * (size_t)i, width, height, env_of_image ? env_of_image[i] : 0);
        voi[i] = r->h_view_of_image[i];
    }
    cudaStream_t st = (cudaStream_t)stream;
    SAS_CUDA(r, cudaMemcpyAsync(r->row_cams, cams, sizeof(SasCamera) * (size_t)(C + n_images), cudaMemcpyHostToDevice, st));
    SAS_CUDA(r, cudaMemcpyAsync(r->view_of_image, voi, sizeof(int) * (size_t)n_images, cudaMemcpyHostToDevice, st));
    SAS_CUDA(r, cudaEventRecord(r->cams_copied, st));
    r->I = n_images; r->C = C; r->W = width; r->H = height;
    r->max_env_index = max_env;
    r->tile_w = (width + SAS_TILE - 1) / SAS_TILE;
    r->tile_h = (height + SAS_TILE - 1) / SAS_TILE;
    r->cameras_set = true;
    // size the grow-only work buffers for this shape
    const size_t T = (size_t)r->tile_w * r->tile_h;
    const size_t n_rec = (size_t)C * r->n_bg + (size_t)n_images * r->n_dyn, RT = (size_t)(C + n_images) * T, IT = (size_t)n_images * T;
    int rc;
    if ((rc = ensure(r, r->rec, n_rec * sizeof(SasRecord)))) return rc;
    if ((rc = ensure(r, r->bins, n_rec * sizeof(uint2)))) return rc;
    if ((rc = ensure(r, r->radii, n_rec * 8))) return rc;
    if ((rc = ensure(r, r->row_tile_count, RT * 4))) return rc;
    if ((rc = ensure(r, r->row_tile_cursor, RT * 4))) return rc;
    if ((rc = ensure(r, r->row_tile_offset, (RT + 1) * 4))) return rc;
    if ((rc = ensure(r, r->img_tile_offset, (IT + 1) * 4))) return rc;
    if ((rc = ensure(r, r->block_sums, 2 * (RT / 1024 + 2) * 4))) return rc;  // one half per scan (they run concurrently)
    if ((rc = ensure(r, r->img_max, (size_t)n_images * 4))) return rc;
    if ((rc = ensure(r, r->dyn_mask, IT * 4))) return rc;
    // work lists (sas_work.cu) and the per-view tile buffers shared by the images of a view
    const size_t VT = (size_t)C * T, VP = (size_t)C * width * height;
    if ((rc = ensure(r, r->sort_list, RT * 4))) return rc;
    if ((rc = ensure(r, r->comp_list, (IT + VT) * 4))) return rc;
    if ((rc = ensure(r, r->view_needed, VT))) return rc;
    if ((rc = ensure(r, r->v_acc, VP * 16))) return rc;
    if ((rc = ensure(r, r->v_depth, VP * 4))) return rc;
    if ((rc = ensure(r, r->v_nc, VP * 4))) return rc;
    return SAS_OK;
}

// Enqueues the device work of one render on st (no host synchronisation, no allocation): the part that is
// captured into a CUDA graph.
static int enqueue_render(sas_renderer *r, const sas_outputs *out, float *d_rgb, float *d_depth, float *d_alpha, uint8_t *d_u8, int *d_nc,
                          uint16_t *d_f16, cudaStream_t st, int *n_launches) {
    const long long cap_src = r->max_isects_src, cap_all = r->max_isects;
    const int I = r->I, C = r->C, W = r->W, H = r->H, tw = r->tile_w, th = r->tile_h, T = tw * th;
    const size_t RT = (size_t)(C + I) * T, IT = (size_t)I * T;
    const bool ns = out->nerfstudio_post != 0;
    int launches = 0;
    SAS_CUDA(r, cudaMemsetAsync(r->row_tile_count.p, 0, RT * 4, st));
    SAS_CUDA(r, cudaMemsetAsync(r->row_tile_cursor.p, 0, RT * 4, st));
    // group reach masks (sas_common.cuh): OR-ed by the link projection; all-ones = "composite every group"
    SAS_CUDA(r, cudaMemsetAsync(r->dyn_mask.p, (sas_variant_flags() & 2) ? 0 : 0xff, IT * 4, st));
    unsigned *dyn_mask = (unsigned *)r->dyn_mask.p;
    SAS_CUDA(r, cudaMemsetAsync(r->counters, 0, sizeof(SasCounters) + sizeof(SasWork), st));  // counters and work lists
    SasWork *work = reinterpret_cast<SasWork *>(r->counters + 1);

    SasRecord *rec = (SasRecord *)r->rec.p;
    uint2 *bins = (uint2 *)r->bins.p, *pairs = (uint2 *)r->pairs.p, *pairs_alt = (uint2 *)r->pairs_alt.p;
    unsigned *rcount = (unsigned *)r->row_tile_count.p, *rcursor = (unsigned *)r->row_tile_cursor.p;
    int *roff = (int *)r->row_tile_offset.p, *ioff = (int *)r->img_tile_offset.p, *merged = (int *)r->merged.p;
    // background once per view, links once per image
    const SasRowSpan span_bg{0, r->n_bg, C, 0, 0};
    const SasRowSpan span_dy{r->n_bg, r->n_dyn, I, C, (long long)C * r->n_bg};

    // stage boundaries: CUDA events when per-stage timing is on, and an NVTX range per stage (host-side enqueue
    // ranges; with SAS_GRAPH=0 they bracket the launches of each stage in a profiler timeline)
    static const char *const kStageNames[SAS_N_STAGES] = {"sas:project", "sas:scan+lists", "sas:emit", "sas:sort+merge", "sas:composite", "sas:post"};
#define SAS_STAGE(i)                                                                   \
    do {                                                                               \
        if (r->stage_timing) SAS_CUDA(r, cudaEventRecord(r->stage_ev[i], st));         \
        if ((i) > 0) nvtxRangePop();                                                   \
        if ((i) < SAS_N_STAGES) nvtxRangePushA(kStageNames[(i) < SAS_N_STAGES ? (i) : 0]); \
    } while (0)
    SAS_STAGE(0);
    // The background span (once per view) and the link span (once per image) touch disjoint rows: with both present
    // the small one runs on a side stream next to the other (projection here, emission below).
    const bool two_spans = span_bg.n > 0 && span_dy.n > 0 && (sas_variant_flags() & 4);
    auto fork_span = [&]() -> cudaError_t {
        cudaError_t e = cudaEventRecord(r->span_ev[0], st);
        return e != cudaSuccess ? e : cudaStreamWaitEvent(r->aux[2], r->span_ev[0], 0);
    };
    auto join_span = [&]() -> cudaError_t {
        cudaError_t e = cudaEventRecord(r->span_ev[1], r->aux[2]);
        return e != cudaSuccess ? e : cudaStreamWaitEvent(st, r->span_ev[1], 0);
    };
    if (two_spans) SAS_CUDA(r, fork_span());
    for (const SasRowSpan &sp : {span_bg, span_dy}) {
        if (sp.n <= 0) continue;
        SAS_CUDA(r, sas_launch_project(r->mean_op, r->quat, r->scale_link, r->sh_planes, r->n_pad, r->sh_degree,
                                       r->L > 0 ? r->link_poses : nullptr, r->L, r->row_cams, sp, W, H, tw, th, rec, bins, (uint2 *)r->radii.p, rcount,
                                       sp.row0 == C ? dyn_mask : nullptr, r->counters, r->extend, (two_spans && sp.row0 == 0) ? r->aux[2] : st));
        launches += 1;
    }
    if (two_spans) SAS_CUDA(r, join_span());
    SAS_STAGE(1);
    if (sas_scan_lists_fused_ok(C, I, T)) {
        SAS_CUDA(r, sas_launch_scan_lists_fused(rcount, r->view_of_image, dyn_mask, C, I, W, H, tw, th, roff, ioff, r->counters, cap_src,
                                                cap_all, work, (unsigned char *)r->view_needed.p, (int *)r->sort_list.p,
                                                (int *)r->comp_list.p, st, &launches));
    } else {
        // The two scans and the work lists only read the counts (and the group masks): three independent chains of
        // 3 + 3 + 4 small kernels.  Forked onto side streams they cost the longest chain instead of the sum
        // (SAS_VARIANT bit 2, as for the sort classes); all three are joined before the emission.
        unsigned *bs_a = (unsigned *)r->block_sums.p, *bs_b = bs_a + (RT / 1024 + 2);
        const bool fork = (sas_variant_flags() & 4) != 0;
        cudaStream_t s_b = fork ? r->aux[0] : st, s_w = fork ? r->aux[1] : st;
        if (fork) {
            SAS_CUDA(r, cudaEventRecord(r->scan_ev[0], st));
            SAS_CUDA(r, cudaStreamWaitEvent(s_b, r->scan_ev[0], 0));
            SAS_CUDA(r, cudaStreamWaitEvent(s_w, r->scan_ev[0], 0));
        }
        SAS_CUDA(r, sas_launch_scan(rcount, (int)RT, nullptr, C, T, roff, bs_a, r->counters, cap_src, st, &launches));
        SAS_CUDA(r, sas_launch_scan(rcount, (int)IT, r->view_of_image, C, T, ioff, bs_b, r->counters, cap_all, s_b, &launches));
        SAS_CUDA(r, sas_launch_work_lists(rcount, r->view_of_image, dyn_mask, C, I, W, H, tw, th, work, (unsigned char *)r->view_needed.p,
                                          (int *)r->sort_list.p, (int *)r->comp_list.p, s_w, &launches));
        if (fork) {
            SAS_CUDA(r, cudaEventRecord(r->scan_ev[1], s_b));
            SAS_CUDA(r, cudaEventRecord(r->scan_ev[2], s_w));
            SAS_CUDA(r, cudaStreamWaitEvent(st, r->scan_ev[1], 0));
            SAS_CUDA(r, cudaStreamWaitEvent(st, r->scan_ev[2], 0));
        }
    }
    SAS_STAGE(2);
    if (two_spans) SAS_CUDA(r, fork_span());
    for (const SasRowSpan &sp : {span_bg, span_dy}) {
        if (sp.n <= 0) continue;
        SAS_CUDA(r, sas_launch_emit(bins, sp, tw, th, roff, rcursor, pairs, cap_src, (two_spans && sp.row0 == 0) ? r->aux[2] : st));
        launches += 1;
    }
    if (two_spans) SAS_CUDA(r, join_span());
    SAS_STAGE(3);
    SAS_CUDA(r, sas_launch_tile_sort(pairs, pairs_alt, roff, (const int *)r->sort_list.p, work, cap_src, r->canon_of, r->n_bg,
                                     (int2 *)r->big_items.p, (int)(r->big_items.bytes / sizeof(int2)), r->big_hint ? 1 : 0, st, r->aux,
                                     r->fork_ev, &launches));
    // Compacted compositor lists (SAS_VARIANT bit 12): worth it when many images share a view -- the reach masks of the
    // background bucket entries are computed once per VIEW entry and save staging work in every image of the view.
    const bool compact = (sas_variant_flags() & 0x1000) && (sas_variant_flags() & 8) && (sas_variant_flags() & 2) && r->n_dyn > 0 && I >= compact_min_ratio() * C &&
                         r->bg_reach.p && r->comp_len.p;
    // bit 13: the view buckets' records materialised with exact reach masks; only the default compositor reads them
    const int vfl = sas_variant_flags();
    const bool use_vrec = compact && (vfl & 0x2000) && (vfl & 0xe71) == 0x241 && r->vrec.p;
    SasRecord *vrec = use_vrec ? (SasRecord *)r->vrec.p : nullptr;
    // bit 14: per-part lists for tiles composited by two CTAs (flag bits 29 / 30 of the codes: ids and positions below 2^29)
    const bool use_parts = use_vrec && (vfl & 0x4000) && r->merged2.p && r->max_isects_src < 0x1ffffff0ll && (long long)r->N < 0x1ffffff0ll;
    int *merged2 = use_parts ? (int *)r->merged2.p : nullptr;
    if (compact) {
        SAS_CUDA(r, sas_launch_bg_reach(pairs, roff, rec, r->n_bg, C, T, tw, cap_src, (unsigned *)r->bg_reach.p, vrec, 0 /* 148 SMs */, st));
        launches += 1;
    }
    SAS_CUDA(r, sas_launch_merge(pairs, roff, ioff, r->view_of_image, C, I, T, merged, cap_src, cap_all,
                                 (sas_variant_flags() & 8) ? 0 : 1, (const int *)r->comp_list.p, work, compact ? (const unsigned *)r->bg_reach.p : nullptr,
                                 compact ? dyn_mask : nullptr, compact ? (int *)r->comp_len.p : nullptr, use_vrec ? 1 : 0, rec, merged2, r->n_bg, r->n_dyn, tw, W, H, st));
    launches += 1;
    SAS_STAGE(4);
    // nerfstudio mode composites over black and blends the background afterwards, like get_outputs does
    const float b0 = ns ? 0.0f : out->background[0], b1 = ns ? 0.0f : out->background[1], b2 = ns ? 0.0f : out->background[2];
    const SasOutputs img_out{d_rgb, d_depth, d_alpha, ns ? nullptr : d_u8, d_nc, d_f16};
    const SasViewBuf view_out{(float *)r->v_acc.p, (float *)r->v_depth.p, (int *)r->v_nc.p};
    SAS_CUDA(r, sas_launch_composite(rec, merged, pairs, ioff, roff, rcount, r->view_of_image, dyn_mask, compact ? (const int *)r->comp_len.p : nullptr, vrec, merged2,
                                     (const int *)r->comp_list.p, work,
                                     r->n_bg, r->n_dyn, C, I, W, H, tw, th, b0, b1, b2, img_out, view_out, cap_src, cap_all,
                                     r->counters, st, &launches));
    SAS_STAGE(5);
    if (ns) {
        SAS_CUDA(r, sas_launch_ns_post(d_rgb, d_depth, d_alpha, d_u8, d_f16, I, W, H, out->background[0], out->background[1],
                                       out->background[2], (unsigned *)r->img_max.p, st, &launches));
    }
    SAS_STAGE(6);
    SAS_CUDA(r, sas_launch_snapshot(r->counters, r->snap, st));  // this render's counters -> ring[seq & 3]
    launches += 1;
#undef SAS_STAGE
    *n_launches = launches;
    return SAS_OK;
}

int sas_render(sas_renderer *r, const sas_outputs *out, void *stream) {
    if (!r || !out) return SAS_EINVAL;
    if (!r->cameras_set) return fail(r, SAS_ESTATE, "sas_render: sas_set_cameras has not been called");
    if (r->n_dyn > 0 && r->n_envs == 0) return fail(r, SAS_ESTATE, "sas_render: scene has links but sas_set_link_poses has not been called");
    SAS_CUDA(r, cudaSetDevice(r->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t P = (size_t)r->I * r->W * r->H;

    float *d_rgb = out->rgb, *d_depth = out->depth, *d_alpha = out->alpha;
    uint8_t *d_u8 = out->rgb_u8;
    int *d_nc = out->n_contrib;
    uint16_t *d_f16 = out->depth_f16;
    if (out->mem != SAS_MEM_HOST && out->mem != SAS_MEM_DEVICE && out->mem != SAS_MEM_HOST_ASYNC)
        return fail(r, SAS_EINVAL, "sas_render: outputs.mem must be SAS_MEM_HOST, SAS_MEM_DEVICE or SAS_MEM_HOST_ASYNC");
    if (r->L > 0 && r->max_env_index >= r->n_envs)
        return fail(r, SAS_ESTATE, "sas_render: an image refers to an env beyond the n_envs given to sas_set_link_poses");
    const bool async_out = out->mem == SAS_MEM_HOST_ASYNC;
    const bool host_out = out->mem == SAS_MEM_HOST || async_out;
    const bool ns = out->nerfstudio_post != 0;
    const int set = async_out ? (int)(r->async_seq & 1u) : 0;  // device staging set of this render
    int rc;
    if (async_out && !r->copy_stream) {
        SAS_CUDA(r, cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking));
        for (cudaEvent_t &ev : r->render_done) SAS_CUDA(r, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (cudaEvent_t &ev : r->copy_done) SAS_CUDA(r, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
    // host outputs: ONE device staging block per set, laid out by sas_packed_layout
    void *const h_ptr[6] = {out->rgb_u8, out->depth_f16, out->depth, out->alpha, out->rgb, out->n_contrib};
    int want = 0;
    PackedLayout lay;
    if (host_out) {
        for (int k = 0; k < 6; k++)
            if (h_ptr[k]) want |= 1 << k;
        lay = packed_layout(P, want);
        if ((rc = ensure(r, r->o_pack[set], lay.total))) return rc;
        unsigned char *base = (unsigned char *)r->o_pack[set].p;
        d_u8 = out->rgb_u8 ? (uint8_t *)(base + lay.off[0]) : nullptr;
        d_f16 = out->depth_f16 ? (uint16_t *)(base + lay.off[1]) : nullptr;
        d_depth = out->depth ? (float *)(base + lay.off[2]) : nullptr;
        d_alpha = out->alpha ? (float *)(base + lay.off[3]) : nullptr;
        d_rgb = out->rgb ? (float *)(base + lay.off[4]) : nullptr;
        d_nc = out->n_contrib ? (int *)(base + lay.off[5]) : nullptr;
    }
    if (ns) {  // the post-processing needs rgb and alpha even when the caller does not ask for them
        if (!d_rgb) { if ((rc = ensure(r, r->o_rgb[0], P * 12))) return rc; d_rgb = (float *)r->o_rgb[0].p; }
        if (!d_alpha) { if ((rc = ensure(r, r->o_alpha[0], P * 4))) return rc; d_alpha = (float *)r->o_alpha[0].p; }
    }
    if ((rc = ensure(r, r->img_max, (size_t)r->I * 4))) return rc;
    {   // (bucket, chunk) items of the buckets beyond shared memory: at most one partial chunk per bucket + the full chunks
        const size_t RTb = (size_t)(r->C + r->I) * r->tile_w * r->tile_h;
        if ((rc = ensure(r, r->big_items, (RTb + (size_t)(r->max_isects_src / SAS_SORT_CAP3) + 16) * sizeof(int2)))) return rc;
        if (r->n_dyn > 0 && r->I >= compact_min_ratio() * r->C) {   // compacted compositor lists (enqueue_render)
            if ((rc = ensure(r, r->bg_reach, (size_t)r->max_isects_src * 4))) return rc;
            if ((rc = ensure(r, r->comp_len, (size_t)r->I * r->tile_w * r->tile_h * 4 * 2))) return rc;  // part 0 | part 1 lengths
            if ((sas_variant_flags() & 0x2000) && (rc = ensure(r, r->vrec, (size_t)r->max_isects_src * sizeof(SasRecord)))) return rc;
            if ((sas_variant_flags() & 0x6000) == 0x6000 && (rc = ensure(r, r->merged2, (size_t)r->max_isects * 4))) return rc;
        }
    }
    // the render before last used this staging set: its D2H copy must have drained before the set is overwritten
    if (async_out && r->async_seq >= 2) SAS_CUDA(r, cudaStreamWaitEvent(st, r->copy_done[set], 0));
    // a blocking host render after async ones reuses staging set 0: wait for the async copy that may still read it
    if (host_out && !async_out && r->async_seq >= 1) SAS_CUDA(r, cudaStreamWaitEvent(st, r->copy_done[0], 0));

    // The step is ~25 launches, all with arguments that do not change from step to step (link poses and
    // cameras live in fixed device buffers), so it is captured once into a CUDA graph and replayed: one
    // launch per step instead of 25 (SAS_GRAPH=0 disables).  Not possible on the legacy / per-thread default
    // streams (capture is not allowed there) or with per-stage timing events.
    if (r->graph_mode < 0) {
        const char *e = getenv("SAS_GRAPH");
        r->graph_mode = e ? atoi(e) : 1;
    }
    const bool can_graph = r->graph_mode != 0 && !r->stage_timing && st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread;
    int launches = 0;
    if (can_graph) {
        sas_renderer::GraphKey key;
        memset(&key, 0, sizeof(key));  // padding bytes take part in the memcmp below
        key.I = r->I; key.C = r->C; key.W = r->W; key.H = r->H; key.ns = ns ? 1 : 0; key.flags = sas_variant_flags();
        key.big = r->big_hint ? 1 : 0;
        for (int k = 0; k < 3; k++) key.bg[k] = out->background[k];
        key.rgb = d_rgb; key.depth = d_depth; key.alpha = d_alpha; key.u8 = d_u8; key.nc = d_nc; key.f16 = d_f16;
        key.max_isects = r->max_isects; key.max_isects_src = r->max_isects_src; key.buf_gen = r->buf_gen;
        sas_renderer::GraphSlot *slot = nullptr;
        for (auto &g : r->graphs)
            if (g.exec && memcmp(&key, &g.key, sizeof(key)) == 0) slot = &g;
        bool seen = false;
        for (int k = 0; k < 2; k++) seen = seen || (r->pending_valid[k] && memcmp(&key, &r->pending_key[k], sizeof(key)) == 0);
        if (!slot && !seen) {
            // first sight of this key (e.g. freshly allocated output tensors every call): launch directly and
            // remember it; capturing + instantiating costs more than one step saves
            memcpy(&r->pending_key[r->pending_next], &key, sizeof(key));
            r->pending_valid[r->pending_next] = true;
            r->pending_next ^= 1;
            if ((rc = enqueue_render(r, out, d_rgb, d_depth, d_alpha, d_u8, d_nc, d_f16, st, &launches))) return rc;
        } else {
            if (!slot) {
                slot = &r->graphs[0];
                if (slot->exec && (!r->graphs[1].exec || r->graphs[1].used < slot->used)) slot = &r->graphs[1];
                if (slot->exec) { cudaGraphExecDestroy(slot->exec); slot->exec = nullptr; }
                cudaGraph_t graph = nullptr;
                SAS_CUDA(r, cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
                rc = enqueue_render(r, out, d_rgb, d_depth, d_alpha, d_u8, d_nc, d_f16, st, &launches);
                cudaError_t ce = cudaStreamEndCapture(st, &graph);
                if (rc != SAS_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
                if (ce != cudaSuccess) return fail(r, SAS_ECUDA, std::string("cudaStreamEndCapture failed: ") + cudaGetErrorString(ce));
                ce = cudaGraphInstantiate(&slot->exec, graph, 0);
                cudaGraphDestroy(graph);
                if (ce != cudaSuccess) {
                    slot->exec = nullptr;
                    return fail(r, SAS_ECUDA, std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(ce));
                }
                memcpy(&slot->key, &key, sizeof(key));
                slot->launches = launches;
            }
            SAS_CUDA(r, cudaGraphLaunch(slot->exec, st));
            slot->used = ++r->graph_clock;
            launches = slot->launches;
        }
    } else {
        if ((rc = enqueue_render(r, out, d_rgb, d_depth, d_alpha, d_u8, d_nc, d_f16, st, &launches))) return rc;
    }
    r->last_launches = launches;
    r->rendered = true;
    const unsigned long long seq = r->render_seq++;   // this render's snapshot slot is seq & 3
    if (host_out) {
        // async: the copies run on the library's copy stream behind an event, so the caller's stream is free for the
        // next render; nothing synchronises here (sas_wait_outputs does)
        cudaStream_t cs = async_out ? r->copy_stream : st;
        if (async_out) {
            SAS_CUDA(r, cudaEventRecord(r->render_done[set], st));
            SAS_CUDA(r, cudaStreamWaitEvent(cs, r->render_done[set], 0));
        }
        // one copy when the caller's host buffers are one block with the staging layout, else one per output
        const unsigned char *d_base = (const unsigned char *)r->o_pack[set].p;
        int first = -1;
        bool packed = want != 0;
        for (int k = 0; k < 6 && packed; k++)
            if (want & (1 << k)) {
                if (first < 0) first = k;
                packed = (unsigned char *)h_ptr[k] - (unsigned char *)h_ptr[first] == (ptrdiff_t)(lay.off[k] - lay.off[first]);
            }
        static const size_t kBytes[6] = {3, 2, 4, 4, 12, 4};
        if (packed) {
            int last = first;
            for (int k = 0; k < 6; k++)
                if (want & (1 << k)) last = k;
            SAS_CUDA(r, cudaMemcpyAsync(h_ptr[first], d_base + lay.off[first], lay.off[last] + P * kBytes[last] - lay.off[first],
                                        cudaMemcpyDeviceToHost, cs));
        } else {
            for (int k = 0; k < 6; k++)
                if (want & (1 << k)) SAS_CUDA(r, cudaMemcpyAsync(h_ptr[k], d_base + lay.off[k], P * kBytes[k], cudaMemcpyDeviceToHost, cs));
        }
        // this render's counter snapshot travels with the observation (the ring slot stays valid for 3 more renders)
        SAS_CUDA(r, cudaMemcpyAsync(&r->h_snap->slot[seq & 3ull], &r->snap->slot[seq & 3ull], sizeof(SasCounters), cudaMemcpyDeviceToHost, cs));
        if (async_out) {
            SAS_CUDA(r, cudaEventRecord(r->copy_done[set], cs));
            r->async_render_seq[set] = seq;
            r->async_seq++;
        } else {
            SAS_CUDA(r, cudaStreamSynchronize(st));
            note_counters(r, r->h_snap->slot[seq & 3ull]);
            if (seq >= r->ovf_floor && r->h_snap->slot[seq & 3ull].overflow) return overflow_error(r, r->h_snap->slot[seq & 3ull], "sas_render");
        }
    }
    return SAS_OK;
}

int sas_wait_outputs(sas_renderer *r, int age) {
    if (!r) return SAS_EINVAL;
    if (age < 0 || age > 1) return fail(r, SAS_EINVAL, "sas_wait_outputs: age must be 0 (latest async render) or 1 (the one before)");
    if (r->async_seq <= (unsigned long long)age) return fail(r, SAS_ESTATE, "sas_wait_outputs: no such SAS_MEM_HOST_ASYNC render was issued");
    SAS_CUDA(r, cudaSetDevice(r->device));
    SAS_CUDA(r, cudaEventSynchronize(r->copy_done[(r->async_seq - 1 - age) & 1u]));
    // the async renders are the last renders issued on this handle unless other modes were mixed in: find the snapshot
    // of this one by its position among ALL renders (async_render_seq remembers the last two)
    const unsigned long long seq = r->async_render_seq[(r->async_seq - 1 - age) & 1u];
    note_counters(r, r->h_snap->slot[seq & 3ull]);
    if (seq >= r->ovf_floor && r->h_snap->slot[seq & 3ull].overflow) return overflow_error(r, r->h_snap->slot[seq & 3ull], "sas_wait_outputs");
    return SAS_OK;
}

int sas_poll_overflow(sas_renderer *r, long long *need_src, long long *need_all, void *stream) {
    if (!r) return SAS_EINVAL;
    SAS_CUDA(r, cudaSetDevice(r->device));
    cudaStream_t st = (cudaStream_t)stream;
    SAS_CUDA(r, cudaMemcpyAsync(r->h_snap, r->snap, sizeof(SasSnapRing), cudaMemcpyDeviceToHost, st));
    SAS_CUDA(r, cudaStreamSynchronize(st));
    const unsigned long long done = r->h_snap->seq;  // renders that have run
    long long ns = 0, na = 0;
    int any = 0;
    for (unsigned long long k = done > 4 ? done - 4 : 0; k < done; k++) {
        if (k < r->ovf_floor) continue;
        const SasCounters &c = r->h_snap->slot[k & 3ull];
        note_counters(r, c);
        if (!c.overflow) continue;
        any = 1;
        if ((long long)c.n_isects_src > ns) ns = (long long)c.n_isects_src;
        if ((long long)c.n_isects > na) na = (long long)c.n_isects;
    }
    r->ovf_floor = done;
    if (need_src) *need_src = ns;
    if (need_all) *need_all = na;
    return any;
}

int sas_set_stage_timing(sas_renderer *r, int enable) {
    if (!r) return SAS_EINVAL;
    SAS_CUDA(r, cudaSetDevice(r->device));
    if (enable)
        for (cudaEvent_t &ev : r->stage_ev)
            if (!ev) SAS_CUDA(r, cudaEventCreate(&ev));
    r->stage_timing = enable != 0;
    return SAS_OK;
}

int sas_get_stage_ms(sas_renderer *r, float *ms_out) {
    if (!r || !ms_out) return SAS_EINVAL;
    if (!r->stage_timing || !r->rendered) return fail(r, SAS_ESTATE, "sas_get_stage_ms: stage timing is off or nothing rendered");
    SAS_CUDA(r, cudaSetDevice(r->device));
    SAS_CUDA(r, cudaEventSynchronize(r->stage_ev[SAS_N_STAGES]));
    for (int i = 0; i < SAS_N_STAGES; i++) SAS_CUDA(r, cudaEventElapsedTime(ms_out + i, r->stage_ev[i], r->stage_ev[i + 1]));
    return SAS_OK;
}

int sas_get_stats(sas_renderer *r, sas_stats *out) {
    if (!r || !out) return SAS_EINVAL;
    SAS_CUDA(r, cudaSetDevice(r->device));
    SAS_CUDA(r, cudaDeviceSynchronize());
    SasCounters c;
    SAS_CUDA(r, cudaMemcpy(&c, r->counters, sizeof(c), cudaMemcpyDeviceToHost));
    note_counters(r, c);
    out->n_isects = (int64_t)c.n_isects;
    out->n_isects_src = (int64_t)c.n_isects_src;
    out->n_visible_src = (int64_t)c.n_visible_src;
    out->n_staged = (int64_t)(c.n_staged / 8);  // the compositor counts per strip of WARPS/8 of a tile
    out->max_tile_list = (int64_t)c.max_tile_list;
    out->overflow = (int32_t)c.overflow;
    out->cap_isects = r->max_isects; out->cap_isects_src = r->max_isects_src;
    out->n_images = r->I; out->n_views = r->C; out->n_gaussians = r->N; out->n_background = r->n_bg;
    out->width = r->W; out->height = r->H; out->tile_w = r->tile_w; out->tile_h = r->tile_h;
    return SAS_OK;
}

int sas_debug_fetch(sas_renderer *r, int what, void *dst_host, size_t bytes) {
    if (!r || !dst_host) return SAS_EINVAL;
    SAS_CUDA(r, cudaSetDevice(r->device));
    SAS_CUDA(r, cudaDeviceSynchronize());
    if (what == SAS_DBG_LINK_POSES) {
        const size_t avail = sizeof(float) * 12 * (size_t)r->n_envs * r->L;
        if (bytes > avail) return fail(r, SAS_EINVAL, "sas_debug_fetch: more bytes requested than available");
        if (bytes) SAS_CUDA(r, cudaMemcpy(dst_host, r->link_poses, bytes, cudaMemcpyDeviceToHost));
        return SAS_OK;
    }
    if (!r->rendered) return fail(r, SAS_ESTATE, "sas_debug_fetch: nothing rendered yet");
    const size_t T = (size_t)r->tile_w * r->tile_h;
    const size_t n_rec = (size_t)r->C * r->n_bg + (size_t)r->I * r->n_dyn, IT = (size_t)r->I * T;
    const void *src = nullptr;
    size_t avail = 0;
    switch (what) {
        case SAS_DBG_RECORDS: src = r->rec.p; avail = n_rec * sizeof(SasRecord); break;
        case SAS_DBG_BINS: src = r->bins.p; avail = n_rec * sizeof(uint2); break;
        case SAS_DBG_RADII: src = r->radii.p; avail = n_rec * 8; break;
        case SAS_DBG_TILE_OFFSETS: src = r->img_tile_offset.p; avail = (IT + 1) * 4; break;
        case SAS_DBG_VIEW_OF_IMAGE: src = r->view_of_image; avail = (size_t)r->I * 4; break;
        case SAS_DBG_SORTED_GIDS: {
            SasCounters c;
            SAS_CUDA(r, cudaMemcpy(&c, r->counters, sizeof(c), cudaMemcpyDeviceToHost));
            long long m = (long long)c.n_isects < r->max_isects ? (long long)c.n_isects : r->max_isects;
            // the render skips the id lists of tiles without link Gaussians (the compositor takes those from the
            // view's bucket); produce them now
            SAS_CUDA(r, sas_launch_merge((const uint2 *)r->pairs.p, (const int *)r->row_tile_offset.p, (const int *)r->img_tile_offset.p,
                                         r->view_of_image, r->C, r->I, (int)T, (int *)r->merged.p, r->max_isects_src, r->max_isects, 1, nullptr, nullptr,
                                         nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0, 0, 0, 0, 0, 0));
            SAS_CUDA(r, cudaDeviceSynchronize());
            src = r->merged.p; avail = (size_t)m * 4;
            break;
        }
        default: return fail(r, SAS_EINVAL, "sas_debug_fetch: unknown item");
    }
    if (bytes > avail) return fail(r, SAS_EINVAL, "sas_debug_fetch: more bytes requested than available");
    if (!bytes) return SAS_OK;
    const bool rows = what == SAS_DBG_RECORDS || what == SAS_DBG_BINS || what == SAS_DBG_RADII;
    if (r->h_canon.empty() || !(rows || what == SAS_DBG_SORTED_GIDS)) {
        SAS_CUDA(r, cudaMemcpy(dst_host, src, bytes, cudaMemcpyDeviceToHost));
        return SAS_OK;
    }
    // the background lives in the internal (Morton) order on the device: hand it out in the canonical order
    std::vector<unsigned char> tmp(avail);
    SAS_CUDA(r, cudaMemcpy(tmp.data(), src, avail, cudaMemcpyDeviceToHost));
    if (rows) {
        const size_t es = what == SAS_DBG_RECORDS ? sizeof(SasRecord) : 8;
        std::vector<unsigned char> view((size_t)r->n_bg * es);
        for (int v = 0; v < r->C; v++) {
            unsigned char *base = tmp.data() + (size_t)v * r->n_bg * es;
            for (int i = 0; i < r->n_bg; i++) memcpy(view.data() + (size_t)r->h_canon[i] * es, base + (size_t)i * es, es);
            memcpy(base, view.data(), view.size());
        }
    } else {
        int *ids = reinterpret_cast<int *>(tmp.data());
        for (size_t k = 0; k < avail / 4; k++)
            if (ids[k] >= 0 && ids[k] < r->n_bg) ids[k] = r->h_canon[ids[k]];
    }
    memcpy(dst_host, tmp.data(), bytes);
    return SAS_OK;
}

}  // extern "C"
