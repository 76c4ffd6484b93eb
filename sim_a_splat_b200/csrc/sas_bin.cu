// sas_bin.cu -- K2 exclusive scans of per-(row,tile) counts with warp-shuffle prefix sums,
// K3 emission of (depth bits, gaussian id) pairs into each (row,tile) bucket, and K5 the merge of
// every image's background bucket (its view's) with its link bucket.
//
// Together with sas_sort.cu this is an MSD radix sort of gsplat's 64-bit key
// (image | tile | fp32 depth bits): the (image,tile) digit is resolved here by a counting
// partition (bucket = digit), the 32 depth bits by 8-bit LSD passes in shared memory
// (sas_sort.cu).  The merged order is bit-identical to a stable sort of the full key with
// emission in ascending Gaussian index: ties on depth are broken by Gaussian id inside a bucket
// (sas_sort.cu) and background-before-link across the two buckets (background ids are the lower
// ones in canonical order).  The scan of the merged counts IS gsplat's isect_offsets.
//
// HBM traffic: scan 8 B per bucket; emit reads 8 B per (row,gaussian), writes 8 B per bucket
// entry; merge reads 8 B per merged intersection and writes 4 B.
#include "sas_common.cuh"

#include <cstdlib>

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ unsigned warp_inclusive_scan(unsigned v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// exclusive scan over the 256 threads of a block; returns this thread's exclusive prefix and the block total
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned *total, unsigned *wsum /*[8]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = warp_inclusive_scan(v);
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    unsigned pre = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
        unsigned s = wsum[w];
        if (w < warp) pre += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return pre + inc - v;
}

// The value scanned at position idx: either the raw per-(row,tile) count, or (merged mode,
// view_of_image != nullptr) the size of image i's merged tile list = its view's background
// bucket + its own link bucket.
__device__ __forceinline__ unsigned scan_value(const unsigned *__restrict__ counts, int idx, const int *__restrict__ view_of_image,
                                               int C, int T) {
    if (view_of_image == nullptr) return counts[idx];
    const int i = idx / T, t = idx - i * T;
    return counts[(size_t)view_of_image[i] * T + t] + counts[(size_t)(C + i) * T + t];
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums(const unsigned *__restrict__ in, int n, const int *__restrict__ view_of_image,
                                                                int C, int T, unsigned *__restrict__ block_sums, SasCounters *counters) {
    __shared__ unsigned wsum[8];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned s = 0, mx = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        unsigned v = (base + k < n) ? scan_value(in, base + k, view_of_image, C, T) : 0u;
        s += v;
        mx = max(mx, v);
    }
    unsigned tot;
    block_exclusive_scan(s, &tot, wsum);
    if (view_of_image != nullptr) {  // longest merged list (diagnostic)
        mx = __reduce_max_sync(0xffffffffu, mx);
        if ((threadIdx.x & 31) == 0 && mx) atomicMax(&counters->max_tile_list, mx);
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

// single block: in-place exclusive scan of block_sums[nb]; writes the grand total and the overflow flag
__global__ void __launch_bounds__(SCAN_THREADS) scan_partials(unsigned *block_sums, int nb, SasCounters *counters, long long capacity,
                                                              int merged) {
    __shared__ unsigned wsum[8];
    unsigned long long carry = 0;
    for (int base = 0; base < nb; base += SCAN_THREADS) {
        int i = base + threadIdx.x;
        unsigned v = (i < nb) ? block_sums[i] : 0u;
        unsigned tot;
        unsigned ex = block_exclusive_scan(v, &tot, wsum);
        if (i < nb) block_sums[i] = (unsigned)(carry + ex);
        carry += tot;
    }
    if (threadIdx.x == 0) {
        if (merged) counters->n_isects = carry; else counters->n_isects_src = carry;
        if (carry > (unsigned long long)capacity) atomicOr(&counters->overflow, merged ? 2u : 1u);
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_final(const unsigned *__restrict__ in, int n, const int *__restrict__ view_of_image, int C,
                                                           int T, const unsigned *__restrict__ block_sums, int *__restrict__ out,
                                                           const SasCounters *counters) {
    __shared__ unsigned wsum[8];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? scan_value(in, base + k, view_of_image, C, T) : 0u;
        s += v[k];
    }
    unsigned tot;
    unsigned ex = block_exclusive_scan(s, &tot, wsum) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = (int)ex;
        ex += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long m = view_of_image != nullptr ? counters->n_isects : counters->n_isects_src;
        out[n] = (int)(m > 0x7fffffffull ? 0x7fffffffull : m);
    }
}

// Emission.  A CTA takes 256 Gaussians of one row.  If one of them covers more than EMIT_FLAT_MIN_TILES
// tiles, the CTA treats its (Gaussian, tile) intersections as ONE flattened list (block scan of the
// per-Gaussian tile counts, then a binary search per list entry), so the threads share the work evenly
// however skewed the footprints are -- a single screen-filling Gaussian (2 700 tiles at 720p) walked by
// one thread used to hold its CTA for milliseconds; otherwise every thread walks its own few tiles.  Slots are reserved per
// CTA in three shared-memory phases:
//   1. every intersection counts into s_cnt[tile];
//   2. one thread per touched tile reserves the CTA's range of that bucket with ONE global atomic
//      (s_base) and clears the counter (a scan over all tiles: a touched-tile list appended through one
//      shared counter serialised ~2000 same-address shared atomics per CTA at 1080p);
//   3. every intersection takes a fresh rank from s_cnt[tile] again and writes its pair at
//      s_base + rank.  The ranks of phase 1 and 3 differ, but any assignment of the reserved slots is
//      fine: slot order inside a bucket is arbitrary, the bucket sort fixes it.
// (One global atomic-with-return per intersection is latency bound: 3.7 % issue utilisation on the
// pusht batch.)  Images with more tiles than fit in shared memory (> SAS_SMEM_TILES_MAX) skip phases
// 1-2 and take a global atomic per intersection, still from the flattened list.
constexpr unsigned EMIT_FLAT_MIN_TILES = 16;

__global__ void __launch_bounds__(256) emit_kernel(const uint2 *__restrict__ bins, SasRowSpan span, int tile_w, int tile_h,
                                                   const int *__restrict__ row_tile_offset, unsigned int *__restrict__ row_tile_cursor,
                                                   uint2 *__restrict__ pairs, long long capacity, int smem_tiles) {
    extern __shared__ unsigned int s_emit[];  // [smem_tiles] counts, [smem_tiles] bases
    unsigned int *s_cnt = s_emit, *s_base = s_emit + smem_tiles;
    __shared__ uint2 s_bn[256];
    __shared__ unsigned int s_pre[257];
    __shared__ unsigned int s_wsum[8];
    for (int t = threadIdx.x; t < smem_tiles; t += 256) s_cnt[t] = 0u;

    const int local = blockIdx.x * 256 + threadIdx.x;
    const int row = blockIdx.y;
    uint2 bn = make_uint2(0u, 0u);
    if (local < span.n) bn = __ldg(bins + (size_t)span.rec0 + (size_t)row * span.n + local);
    s_bn[threadIdx.x] = bn;
    // bn.x == 0 (culled) gives a 0 x 0 rectangle: a visible Gaussian always has x1 >= 1
    const int x0 = bn.x & 0xff, y0 = (bn.x >> 8) & 0xff, x1 = (bn.x >> 16) & 0xff, y1 = bn.x >> 24;
    const unsigned ntile = (unsigned)((x1 - x0) * (y1 - y0));
    // CTA-uniform choice (also the barrier that publishes s_cnt / s_bn): flattened list only if some Gaussian
    // of the CTA covers many tiles; otherwise every thread walks its own small rectangle (no searches)
    const bool flat = __syncthreads_or(ntile > EMIT_FLAT_MIN_TILES);
    unsigned total = 0;
    if (flat) {
        const unsigned pre = block_exclusive_scan(ntile, &total, s_wsum);
        s_pre[threadIdx.x] = pre;
        if (threadIdx.x == 0) s_pre[256] = total;
        __syncthreads();
    }
    const size_t tbase = (size_t)(span.row0 + row) * tile_w * tile_h;
    const unsigned gid0 = (unsigned)(span.g0 + blockIdx.x * 256);  // canonical index of the CTA's first Gaussian

    // f(g, t) for every intersection this thread is responsible for: Gaussian g of the CTA, tile t of the image
    auto for_each_entry = [&](auto f) {
        if (flat) {
            for (unsigned k = threadIdx.x; k < total; k += 256) {
                int lo = 0, hi = 256;  // largest g with s_pre[g] <= k
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (s_pre[mid] <= k) lo = mid; else hi = mid;
                }
                const unsigned bx = s_bn[lo].x, r = k - s_pre[lo];
                const int gx0 = bx & 0xff, gy0 = (bx >> 8) & 0xff, rw = (int)((bx >> 16) & 0xff) - gx0;
                const int dy = (int)r / rw;
                f(lo, (gy0 + dy) * tile_w + gx0 + ((int)r - dy * rw));
            }
        } else {
            for (int y = y0; y < y1; y++)
                for (int x = x0; x < x1; x++) f((int)threadIdx.x, y * tile_w + x);
        }
    };

    if (smem_tiles) {
        for_each_entry([&](int, int t) { atomicAdd(&s_cnt[t], 1u); });
        __syncthreads();
        for (int t = threadIdx.x; t < smem_tiles; t += 256) {
            const unsigned c = s_cnt[t];
            if (c) {
                s_base[t] = (unsigned)__ldg(row_tile_offset + tbase + t) + atomicAdd(row_tile_cursor + tbase + t, c);
                s_cnt[t] = 0u;
            }
        }
        __syncthreads();
    }
    for_each_entry([&](int g, int t) {
        const long long pos = smem_tiles ? (long long)s_base[t] + (long long)atomicAdd(&s_cnt[t], 1u)
                                         : (long long)__ldg(row_tile_offset + tbase + t) + (long long)atomicAdd(row_tile_cursor + tbase + t, 1u);
        if (pos < capacity) pairs[pos] = make_uint2(s_bn[g].y, gid0 + (unsigned)g);
    });
}

// K5: merged[image,tile] = stable merge of the view's sorted background bucket and the image's
// sorted link bucket, by depth bits, background first on ties (background ids are the lower ones in
// canonical order, so this is the order a stable sort of gsplat's keys gives).
//
// Merge-path: one CTA per (image, tile).  The output is cut into chunks of 1024 ids which the CTA
// produces in order: per chunk the next (up to) 1024 entries of BOTH lists are staged in shared
// memory (where the chunk starts in either list is known from the previous chunk), every thread
// finds the split of its own 8-id diagonal with a search in shared memory and merges 8 ids
// sequentially, and the ids leave through shared memory in coalesced rows.  Tiles whose link
// bucket is empty are skipped (the compositor takes them from the view's bucket) unless fill_static.
constexpr int MERGE_THREADS = 128;
constexpr int MERGE_VT = 8;
constexpr int MERGE_CHUNK = MERGE_THREADS * MERGE_VT;

// Ablation knob (SAS_COMPACT_MAX, default 32 = every dynamic tile): compact only the tiles whose link Gaussians reach at
// most this many of the 32 pixel groups.  Measured (profiles/r02b_compact_lists.txt): most of the saving is in the tiles
// with 29..32 groups reached -- they are also the longest lists -- so the gate stays open.
static int compact_max_groups() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SAS_COMPACT_MAX");
        v = e ? atoi(e) : 32;
        if (v < 0 || v > 32) v = 32;
    }
    return v;
}

// number of background entries among the first d merged entries
template <typename KA, typename KB>
__device__ __forceinline__ int merge_path(int d, int na, int nb, KA key_a, KB key_b) {
    int lo = max(0, d - nb), hi = min(d, na);
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (key_a(mid) <= key_b(d - 1 - mid)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// bg_reach_kernel: for every entry of the VIEW buckets (background Gaussians, sorted) the 32-bit mask of the tile's 4x2
// groups its culling box reaches -- the same mask the compositor computes when it stages the record, but it does not
// depend on the env, so it is computed once per view entry.  The merge uses it to COMPACT the lists it hands to the
// compositor: a background record that reaches none of the groups an image tile's link Gaussians can reach cannot touch
// a pixel composited from that list (the other groups are copied from the view tile), so it is left out -- 28 % of the
// records of the pusht batch's dynamic tiles.
//
// vrec != nullptr (SAS_VARIANT bit 13): the entry's record is also copied into vrec[k] -- the view bucket's records in
// depth order, contiguous -- with the reach mask in b.z (b.w = -1 marks it) so that the compositor neither gathers the
// background records from the view's 5 MB record array nor computes their masks; and because the mask is now computed
// once per VIEW entry and not once per image, it can afford to be exact: of the groups the culling box reaches only those
// stay where some pixel centre passes the compositor's own range test (its own sigma expression, sas_refine_reach).
__device__ __forceinline__ unsigned sas_refine_reach(unsigned m, const float4 A, const float4 B, int tx, int ty) {  // m: the groups this LANE tests
    // alpha = min(0.999, opacity * exp(-sigma)) >= 1/255 needs sigma <= ln(255 opacity); + 0.01 covers logf and the
    // compositor's polynomial exp.  A negative limit: the Gaussian cannot contribute anywhere.
    const float lim = fminf(SAS_SIGMA_SKIP, logf(255.0f * B.y) + 0.01f);
    if (!(lim >= 0.0f)) return 0u;
    const unsigned ulim = __float_as_uint(lim);
    // centre of the tile's first pixel; every pixel centre below is this plus small integers: exact, and the same float
    // the compositor forms as (float)px + 0.5f
    const float X0 = (float)(tx * SAS_TILE_PX) + 0.5f, Y0 = (float)(ty * SAS_TILE_PX) + 0.5f;
    // exactly blend_one's sigma (sas_composite.cu) and its one-compare range test
    auto passes = [&](float pxf, float pyf) {
        const float dx = A.x - pxf, dy = A.y - pyf;
        const float sigma = fmaf(fmaf(A.z, dx, A.w * dy), dx, (B.x * dy) * dy);
        return __float_as_uint(sigma) <= ulim;
    };
    unsigned out = 0u;
    while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1u;
        const int cx = ((bit >> 2) & 1) * 2 + (bit & 1), ry = (bit >> 3) * 2 + ((bit >> 1) & 1);
        const float gx = X0 + (float)(4 * cx), gy = Y0 + (float)(2 * ry);  // centre of the group's first pixel
        // the pixel nearest to the mean first: most reached groups pass there (claiming a hit is always safe, only a miss
        // must be proven -- by all eight pixels)
        const float nx = fminf(fmaxf(rintf(A.x - gx), 0.0f), 3.0f), ny = fminf(fmaxf(rintf(A.y - gy), 0.0f), 1.0f);
        bool hit = passes(gx + nx, gy + ny);
        if (!hit) {
#pragma unroll 1
            for (int q = 0; q < 8 && !hit; q++) hit = passes(gx + (float)(q & 3), gy + (float)(q >> 2));
        }
        if (hit) out |= 1u << bit;
    }
    return out;
}

__global__ void __launch_bounds__(256) bg_reach_kernel(const uint2 *__restrict__ pairs, const int *__restrict__ row_tile_offset,
                                                       const SasRecord *__restrict__ rec, int n_bg, int CT, int T, int tile_w, long long cap_src,
                                                       unsigned *__restrict__ bg_reach, SasRecord *__restrict__ vrec) {
    // The view buckets are the first C*T buckets, i.e. entries [0, row_tile_offset[C*T]) of the pair array.  Flat over the
    // entries (a grid-stride loop: the count lives on the device), every thread finds its entry's bucket by a binary
    // search over the offsets -- one CTA per bucket left the 20 000-entry buckets of a 720p view to one CTA each.
    // Four lanes share an entry (the exact test of a mask's groups is a serial chain: up to 32 groups x 9 sigma
    // evaluations for a thin Gaussian that crosses the tile): lane j tests the groups whose bit index is j mod 4.
    long long n_view = row_tile_offset[CT];
    if (n_view > cap_src) n_view = cap_src;
    const int sub = threadIdx.x & 3;
    const long long n_iter = (n_view + (long long)gridDim.x * (blockDim.x >> 2) - 1) / ((long long)gridDim.x * (blockDim.x >> 2));
    for (long long it = 0; it < n_iter; it++) {  // uniform trip count: the shuffles below need the whole warp
        const long long k = (it * gridDim.x + blockIdx.x) * (blockDim.x >> 2) + (threadIdx.x >> 2);
        const bool live = k < n_view;
        unsigned m = 0u;
        float4 A = make_float4(0.f, 0.f, 0.f, 0.f), B = A;
        const SasRecord *r = rec;
        if (live) {
            int lo = 0, hi = CT;  // last bucket whose offset is <= k (empty buckets share an offset: the last one holds k)
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(row_tile_offset + mid) <= k) lo = mid; else hi = mid;
            }
            const int v = lo / T, t = lo - v * T;
            const int ty = t / tile_w, tx = t - ty * tile_w;
            r = rec + (size_t)v * n_bg + pairs[k].y;
            A = r->a; B = r->b;
            m = sas_group_reach_mask(sas_group_reach(make_float4(A.x - B.z, A.x + B.z, A.y - B.w, A.y + B.w)), tx, ty);
            if (vrec) m = sas_refine_reach(m & (0x11111111u << sub), A, B, tx, ty);
        }
        if (vrec) {  // OR over the entry's four lanes
            m |= __shfl_xor_sync(0xffffffffu, m, 1);
            m |= __shfl_xor_sync(0xffffffffu, m, 2);
            if (live) {  // lanes 0..2 write the three float4 of the record
                if (sub == 0) vrec[k].a = A;
                else if (sub == 1) vrec[k].b = make_float4(B.x, B.y, __uint_as_float(m), -1.0f);
                else if (sub == 2) vrec[k].c = r->c;
            }
        }
        if (live && sub == 0) bg_reach[k] = m;
    }
}

// bg_reach / tile_dyn_mask / comp_len (all or none): compact the merged list of every image tile to the records the
// compositor needs (see bg_reach_kernel) and store its length; the ids stay in merged order at the tile's offset.
//
// PARTS (SAS_VARIANT bit 14, with bits 12 + 13): the compositor cuts a tile with more than 16 link-reachable groups into two
// CTAs ("parts": the first 16 packed groups, the rest), and each used to stage the tile's whole list although a record
// reaches ~1.5 of a CTA's 16 groups.  Here the merge writes TWO lists, `merged` for part 0 and mp.merged2 for part 1 (same
// offsets; comp_len[item] and comp_len[I*T + item] entries): a background entry goes to the lists of the parts its exact
// reach mask (reach pass) touches, a link entry to both.  An entry absent from a part's list could not have been walked by
// any of its groups, so nothing composited changes.  Codes: bit 31 background (position in vrec), bits 30 / 29 "in part 0 / 1"
// while in shared memory (cleared on the way out), ids and positions below 2^29.
struct SasMergeParts {
    const SasRecord *rec;   // all records (link records of image i: rec + C n_bg + i n_dyn - n_bg, indexed by id)
    unsigned *reach;        // the reach-mask array of the reach pass (one word per pair position): the link buckets' part is free
    int *merged2;           // part 1 lists
    int n_bg, n_dyn, tile_w, W, H;
};

template <bool PARTS>
__global__ void __launch_bounds__(MERGE_THREADS, 12) merge_kernel(const uint2 *__restrict__ pairs, const int *__restrict__ row_tile_offset,
                                                              const int *__restrict__ img_tile_offset, const int *__restrict__ view_of_image,
                                                              int C, int T, int *__restrict__ merged, long long cap_src, long long cap_all,
                                                              int fill_static, const int *__restrict__ comp_list,
                                                              const SasWork *__restrict__ work, int I, const unsigned *__restrict__ bg_reach,
                                                              const unsigned *__restrict__ tile_dyn_mask, int *__restrict__ comp_len,
                                                              int compact_max_groups, int vrec_codes, SasMergeParts mp) {
    __shared__ uint2 sa[MERGE_CHUNK], sb[MERGE_CHUNK];
    int *const s_out = reinterpret_cast<int *>(sa);  // the merged ids leave through sa once every thread holds its ids in registers
    int *const s_out1 = s_out + MERGE_CHUNK;          // PARTS: part 1's ids, the upper half of sa
    __shared__ int s_used;
    __shared__ int s_wsum[MERGE_THREADS / 32];
    int t, i;
    if (comp_list) {
        // render path: CTA k takes entry k of the compositor's work list -- exactly the (image, tile) lists that need
        // merging, longest first, so the 9 000-entry tiles start at once instead of finishing alone at the end
        if (blockIdx.x >= work->comp_total) return;
        const long long item = comp_list[blockIdx.x];
        if (item >= (long long)I * T) return;  // a view tile: its bucket is composited as it is
        i = (int)(item / T); t = (int)(item - (long long)i * T);
    } else {
        t = blockIdx.x; i = blockIdx.y;
    }
    const size_t bb = (size_t)view_of_image[i] * T + t, db = (size_t)(C + i) * T + t;
    long long b0 = row_tile_offset[bb], b1 = row_tile_offset[bb + 1];
    long long d0 = row_tile_offset[db], d1 = row_tile_offset[db + 1];
    const long long o0 = img_tile_offset[(size_t)i * T + t];
    if (b0 < 0 || b1 < b0 || d0 < 0 || d1 < d0 || o0 < 0) return;
    if (b1 > cap_src) b1 = cap_src;
    if (d1 > cap_src) d1 = cap_src;
    const int na = b1 > b0 ? (int)(b1 - b0) : 0, nb = d1 > d0 ? (int)(d1 - d0) : 0;
    if (nb == 0 && !fill_static) return;  // composited from the view's bucket; ids only needed by debug fetches
    const uint2 *bgp = pairs + b0, *dyp = pairs + d0;
    const int n = na + nb;
    const bool has_len = comp_len != nullptr && bg_reach != nullptr && tile_dyn_mask != nullptr;
    if (nb == 0 || (na == 0 && !PARTS)) {  // only one list (a static tile, or no background here): plain copy
        const uint2 *src = nb == 0 ? bgp : dyp;
        for (int k = threadIdx.x; k < n; k += MERGE_THREADS) {
            const long long pos = o0 + k;
            if (pos < cap_all) merged[pos] = (int)src[k].y;
        }
        if (has_len && threadIdx.x == 0) comp_len[(size_t)i * T + t] = n;
        return;
    }
    const unsigned dyn = has_len ? tile_dyn_mask[(size_t)i * T + t] : 0xffffffffu;
    // a tile whose link Gaussians reach (nearly) every group drops (nearly) nothing: plain merge, full length
    const bool compact = has_len && (PARTS || __popc(dyn) <= compact_max_groups);
    const unsigned *brp = compact ? bg_reach + b0 : nullptr;
    long long o_run = o0;  // compaction: where the next chunk's kept ids go
    long long o_run1 = o0;  // PARTS: the same for part 1's list (in mp.merged2)
    // PARTS: the groups of part 0 (the first 16 set bits of the tile's packed-group mask, exactly the compositor's cut)
    // and of part 1 (the rest; empty for a tile of one CTA)
    unsigned pm0 = 0u, pm1 = 0u;
    if (PARTS) {
        const unsigned pm = dyn & sas_inside_group_mask(t, mp.tile_w, mp.W, mp.H);
        pm0 = pm;
        if (__popc(pm) > 16) {
            const unsigned b16 = __fns(pm, 0u, 16);
            pm0 = pm & ((2u << b16) - 1u);
            pm1 = pm & ~pm0;
        }
    }
    // Streaming merge: the CTA walks the tile's chunks in order, so it knows where chunk c starts in both lists once
    // chunk c-1 is merged -- no merge-path split search over global memory (a dozen dependent L2 round trips per tile,
    // more than half of a CTA's life in round 1's kernel).  A chunk of cn outputs takes at most cn entries of either
    // list: both windows are loaded (coalesced; the unused tail is loaded again by the next chunk) and the merge path
    // is resolved in shared memory only.
    const int n_chunks = (n + MERGE_CHUNK - 1) / MERGE_CHUNK;
    int a_pos = 0, b_pos = 0;  // entries of either list consumed so far
    for (int c = 0; c < n_chunks; c++) {
        const int cn = min(MERGE_CHUNK, n - c * MERGE_CHUNK);
        const int wa = min(cn, na - a_pos), wb = min(cn, nb - b_pos);
        __syncthreads();  // the previous chunk's readers of sa / sb / s_out / s_used are done
        for (int k = threadIdx.x; k < wa; k += MERGE_THREADS) {
            uint2 e = bgp[a_pos + k];
            // a background record that reaches none of the groups the link Gaussians can reach is not needed: id -1.
            // vrec_codes: the others are named by their position in the view bucket's contiguous records (top bit set;
            // positions and ids are < 2^31 - 16) instead of their id -- also in a tile the SAS_COMPACT_MAX gate keeps
            // uncompacted: the compositor reads every background record of this render from vrec
            if (PARTS) {
                const unsigned m = brp[a_pos + k];
                const unsigned f = ((m & pm0) ? 0x40000000u : 0u) | ((m & pm1) ? 0x20000000u : 0u);
                e.y = f ? (0x80000000u | f | (unsigned)(b0 + a_pos + k)) : 0xffffffffu;
            } else if (compact && (brp[a_pos + k] & dyn) == 0u) e.y = 0xffffffffu;
            else if (vrec_codes) e.y = 0x80000000u | (unsigned)(b0 + a_pos + k);
            sa[k] = e;
        }
        for (int k = threadIdx.x; k < wb; k += MERGE_THREADS) {
            uint2 e = dyp[b_pos + k];
            // link entries go to both lists: their masks would have to be gathered from the link records (measured: 660 MB
            // of L2 traffic and a latency-bound merge, 0.09 -> 0.24 ms, against 0.04 ms saved in the compositor)
            if (PARTS) e.y |= pm1 ? 0x60000000u : 0x40000000u;
            sb[k] = e;
        }
        __syncthreads();
        const int d = threadIdx.x * MERGE_VT;
        int ids[MERGE_VT];
        unsigned keep = 0u;  // bit v: output d + v exists and is kept (PARTS: in part 0's list)
        unsigned keep1 = 0u;  // PARTS: ... in part 1's list
        if (d < cn) {
            int ia = merge_path(d, wa, wb, [&](int k) { return sa[k].x; }, [&](int k) { return sb[k].x; });
            int ib = d - ia;
#pragma unroll
            for (int v = 0; v < MERGE_VT; v++) {
                ids[v] = 0;
                if (d + v < cn) {
                    const bool take_a = ib >= wb || (ia < wa && sa[ia].x <= sb[ib].x);
                    ids[v] = (int)(take_a ? sa[ia].y : sb[ib].y);
                    if (PARTS) {
                        const unsigned cd = (unsigned)ids[v];
                        if (cd != 0xffffffffu) {
                            keep |= ((cd >> 30) & 1u) << v;
                            keep1 |= ((cd >> 29) & 1u) << v;
                            ids[v] = (int)(cd & 0x9fffffffu);
                        }
                    } else if (ids[v] != -1) keep |= 1u << v;
                    if (take_a) ia++; else ib++;
                }
            }
            if (d + MERGE_VT >= cn) s_used = ia;  // the thread that merges the chunk's last output: background entries used
        }
        int cn_out = cn, cn_out1 = 0;
        if (!compact) {
            __syncthreads();  // s_out aliases sa
#pragma unroll
            for (int v = 0; v < MERGE_VT; v++)
                if (keep & (1u << v)) s_out[d + v] = ids[v];
        } else {  // exclusive scan of the kept counts over the CTA, then every thread writes its kept ids densely
            // PARTS: both lists' counts in one word (part 1's in the upper half; a chunk holds 1 024 entries)
            const int cnt = __popc(keep) | (PARTS ? __popc(keep1) << 16 : 0), lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            int inc = cnt;
#pragma unroll
            for (int s_ = 1; s_ < 32; s_ <<= 1) {
                const int tt = __shfl_up_sync(0xffffffffu, inc, s_);
                if (lane >= s_) inc += tt;
            }
            if (lane == 31) s_wsum[warp] = inc;
            __syncthreads();
            int pre = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < MERGE_THREADS / 32; w++) {
                const int sw = s_wsum[w];
                if (w < warp) pre += sw;
                tot += sw;
            }
            int pos = (pre + inc - cnt) & 0xffff, pos1 = (pre + inc - cnt) >> 16;
#pragma unroll
            for (int v = 0; v < MERGE_VT; v++) {
                if (keep & (1u << v)) s_out[pos++] = ids[v];
                if (PARTS && (keep1 & (1u << v))) s_out1[pos1++] = ids[v];
            }
            cn_out = tot & 0xffff;
            cn_out1 = tot >> 16;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < cn_out; k += MERGE_THREADS) {  // coalesced write-out
            const long long pos = o_run + k;
            if (pos < cap_all) merged[pos] = s_out[k];
        }
        o_run += cn_out;
        if (PARTS) {
            for (int k = threadIdx.x; k < cn_out1; k += MERGE_THREADS) {
                const long long pos = o_run1 + k;
                if (pos < cap_all) mp.merged2[pos] = s_out1[k];
            }
            o_run1 += cn_out1;
        }
        const int used_a = s_used;
        a_pos += used_a;
        b_pos += cn - used_a;
    }
    if (has_len && threadIdx.x == 0) {
        comp_len[(size_t)i * T + t] = (int)(o_run - o0);
        if (PARTS) comp_len[(size_t)I * T + (size_t)i * T + t] = (int)(o_run1 - o0);
    }
}

}  // namespace

cudaError_t sas_launch_scan(const unsigned int *counts, int n, const int *view_of_image, int C, int T, int *offsets,
                            unsigned int *block_sums, SasCounters *counters, long long capacity, cudaStream_t st, int *n_launches) {
    const int nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_block_sums<<<nb, SCAN_THREADS, 0, st>>>(counts, n, view_of_image, C, T, block_sums, counters);
    scan_partials<<<1, SCAN_THREADS, 0, st>>>(block_sums, nb, counters, capacity, view_of_image != nullptr ? 1 : 0);
    scan_final<<<nb, SCAN_THREADS, 0, st>>>(counts, n, view_of_image, C, T, block_sums, offsets, counters);
    if (n_launches) *n_launches += 3;
    return cudaGetLastError();
}

cudaError_t sas_launch_emit(const uint2 *bins, SasRowSpan span, int tile_w, int tile_h, const int *row_tile_offset,
                            unsigned int *row_tile_cursor, uint2 *pairs, long long capacity, cudaStream_t st) {
    if (span.n <= 0 || span.rows <= 0) return cudaSuccess;
    dim3 grid((span.n + 255) / 256, span.rows);
    const int T = tile_w * tile_h;
    const int smem_tiles = T <= SAS_SMEM_TILES_MAX ? T : 0;
    emit_kernel<<<grid, 256, 2 * (size_t)smem_tiles * sizeof(unsigned int), st>>>(bins, span, tile_w, tile_h, row_tile_offset,
                                                                                 row_tile_cursor, pairs, capacity, smem_tiles);
    return cudaGetLastError();
}

cudaError_t sas_bin_init() {
    return cudaFuncSetAttribute(emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * SAS_SMEM_TILES_MAX * (int)sizeof(unsigned int));
}

cudaError_t sas_launch_bg_reach(const uint2 *pairs, const int *row_tile_offset, const SasRecord *rec, int n_bg, int C, int T, int tile_w,
                                long long cap_src, unsigned *bg_reach, SasRecord *vrec, int n_sms, cudaStream_t st) {
    if (C <= 0 || T <= 0) return cudaSuccess;
    bg_reach_kernel<<<(n_sms > 0 ? n_sms : 148) * 16, 256, 0, st>>>(pairs, row_tile_offset, rec, n_bg, C * T, T, tile_w, cap_src, bg_reach, vrec);
    return cudaGetLastError();
}

cudaError_t sas_launch_merge(const uint2 *pairs, const int *row_tile_offset, const int *img_tile_offset, const int *view_of_image,
                             int C, int I, int T, int *merged_gid, long long cap_src, long long cap_all, int fill_static,
                             const int *comp_list, const SasWork *work, const unsigned *bg_reach, const unsigned *tile_dyn_mask,
                             int *comp_len, int vrec_codes, const SasRecord *rec, int *merged2, int n_bg, int n_dyn, int tile_w, int W,
                             int H, cudaStream_t st) {
    static int lpt = -1;  // SAS_MERGE_LPT=0: one CTA per (tile, image) in grid order, as for debug fetches
    if (lpt < 0) { const char *e = getenv("SAS_MERGE_LPT"); lpt = e ? atoi(e) : 1; }
    const SasMergeParts mp{rec, const_cast<unsigned *>(bg_reach), merged2, n_bg, n_dyn, tile_w, W, H};
    if (comp_list && work && !fill_static && lpt) {
        const unsigned grid = (unsigned)((long long)(I + C) * T);
        if (merged2 && rec && comp_len && bg_reach && tile_dyn_mask && vrec_codes)
            merge_kernel<true><<<grid, MERGE_THREADS, 0, st>>>(pairs, row_tile_offset, img_tile_offset, view_of_image, C, T, merged_gid, cap_src,
                                                               cap_all, 0, comp_list, work, I, bg_reach, tile_dyn_mask, comp_len,
                                                               compact_max_groups(), vrec_codes, mp);
        else
            merge_kernel<false><<<grid, MERGE_THREADS, 0, st>>>(pairs, row_tile_offset, img_tile_offset, view_of_image, C, T, merged_gid, cap_src,
                                                                cap_all, 0, comp_list, work, I, bg_reach, tile_dyn_mask, comp_len,
                                                                compact_max_groups(), vrec_codes, mp);
    } else {
        dim3 grid(T, I);
        merge_kernel<false><<<grid, MERGE_THREADS, 0, st>>>(pairs, row_tile_offset, img_tile_offset, view_of_image, C, T, merged_gid, cap_src,
                                                            cap_all, fill_static, nullptr, nullptr, I, nullptr, nullptr, nullptr, 32, 0, mp);
    }
    return cudaGetLastError();
}
