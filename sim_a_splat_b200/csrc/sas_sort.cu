// sas_sort.cu -- K4: per-(image,tile) bucket sort by (fp32 depth bits, Gaussian id).
//
// Second level of the MSD radix sort of gsplat's (image | tile | depth) key (first level: the
// counting partition in sas_bin.cu).  One CTA per bucket: the bucket is staged into shared
// memory, sorted there with four stable 8-bit LSD passes over the depth bits (per-warp
// histograms, a warp-shuffle scan across digits, and a stable multi-split that ranks equal
// digits inside a warp with match.any + popc), equal-depth runs are then ordered by Gaussian id
// (which is what a stable sort with ascending-Gaussian emission produces in the reference), and
// the sorted Gaussian ids are written back.  Passes whose digit is uniform over the bucket (the
// exponent byte, usually) are skipped.
//
// Buckets are dispatched by size class (shared-memory footprint = occupancy): <=2048, <=6144,
// <=12288 entries in shared memory; larger buckets run the same passes directly on global memory
// (ping-pong with a second pair buffer) in one CTA.
//
// HBM traffic: 8 B read + 4 B written per intersection.
#include "sas_common.cuh"

namespace {

constexpr int SORT_THREADS = 512;
constexpr int SORT_WARPS = SORT_THREADS / 32;

struct alignas(16) SortShared {
    unsigned hist[SORT_WARPS][256];
    unsigned tot[256];
    unsigned wsum[8];
    int uniform;
};

// One stable LSD pass src -> dst on bits [shift, shift+8) of .x.  Returns false (and leaves dst
// untouched) if every key has the same digit.
__device__ bool radix_pass(const uint2 *src, uint2 *dst, int n, int shift, SortShared &sh) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < SORT_WARPS * 256; i += SORT_THREADS) (&sh.hist[0][0])[i] = 0u;
    if (tid == 0) sh.uniform = 0;
    __syncthreads();
    // contiguous segment per warp, in rounds of 32 consecutive keys (keeps the pass stable)
    const int seg = (((n + SORT_WARPS - 1) / SORT_WARPS) + 31) & ~31;
    const int lo = min(n, warp * seg), hi = min(n, lo + seg);
    for (int i = lo + lane; i < hi; i += 32) atomicAdd(&sh.hist[warp][(src[i].x >> shift) & 0xffu], 1u);
    __syncthreads();
    if (tid < 256) {
        unsigned run = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            unsigned t = sh.hist[w][tid];
            sh.hist[w][tid] = run;
            run += t;
        }
        if (run == (unsigned)n) sh.uniform = 1;
        // exclusive scan of tot over the 256 digits (8 warps)
        unsigned inc = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) sh.wsum[warp] = inc;
        sh.tot[tid] = inc - run;  // exclusive within the warp
    }
    __syncthreads();
    const int uniform = sh.uniform;  // read before anyone can re-enter and reset it
    __syncthreads();
    if (uniform) return false;
    if (tid < 256) {
        unsigned pre = 0;
        for (int w = 0; w < warp; w++) pre += sh.wsum[w];
        const unsigned base = sh.tot[tid] + pre;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) sh.hist[w][tid] += base;
    }
    __syncthreads();
    // stable multi-split: hist[warp][digit] is the running absolute position
    for (int i0 = lo; i0 < hi; i0 += 32) {
        const int i = i0 + lane;
        const bool active = i < hi;
        const unsigned amask = __ballot_sync(0xffffffffu, active);
        if (active) {
            const uint2 kv = src[i];
            const unsigned d = (kv.x >> shift) & 0xffu;
            const unsigned peers = __match_any_sync(amask, d);
            const int leader = __ffs(peers) - 1;
            const unsigned rank = __popc(peers & ((1u << lane) - 1u));
            unsigned base = 0;
            if (lane == leader) {
                base = sh.hist[warp][d];
                sh.hist[warp][d] = base + __popc(peers);
            }
            base = __shfl_sync(peers, base, leader);
            dst[base + rank] = kv;
        }
        __syncwarp();
    }
    __syncthreads();
    return true;
}

// Equal-depth runs -> ascending Gaussian id (runs are disjoint; the thread at a run's start sorts it).
__device__ void fix_ties(uint2 *a, int n) {
    for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
        const unsigned k = a[i].x;
        if ((i == 0 || a[i - 1].x != k) && i + 1 < n && a[i + 1].x == k) {
            int j = i + 1;
            while (j < n && a[j].x == k) j++;
            for (int p = i + 1; p < j; p++) {
                const unsigned v = a[p].y;
                int q = p - 1;
                while (q >= i && a[q].y > v) { a[q + 1].y = a[q].y; q--; }
                a[q + 1].y = v;
            }
        }
    }
    __syncthreads();
}

template <bool IN_GLOBAL>
__global__ void __launch_bounds__(SORT_THREADS) tile_sort_kernel(uint2 *pairs, uint2 *pairs_alt,
                                                                 const int *__restrict__ tile_offset, int *__restrict__ sorted_gid,
                                                                 int lo_size, int hi_size, long long max_isects) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortShared &sh = *reinterpret_cast<SortShared *>(smem_raw);
    const int b = blockIdx.x;
    long long off = tile_offset[b], end = tile_offset[b + 1];
    if (off < 0 || end < off) return;
    if (end > max_isects) end = max_isects;
    const long long nn = end - off;
    if (nn < (long long)lo_size || nn >= (long long)hi_size || nn <= 0) return;
    const int n = (int)nn;
    uint2 *src, *dst;
    if (IN_GLOBAL) {
        src = pairs + off;
        dst = pairs_alt + off;
    } else {
        src = reinterpret_cast<uint2 *>(smem_raw + sizeof(SortShared));
        dst = src + (hi_size - 1);
        for (int i = threadIdx.x; i < n; i += SORT_THREADS) src[i] = pairs[off + i];
    }
    __syncthreads();
    if (n > 1) {
#pragma unroll 1
        for (int shift = 0; shift < 32; shift += 8) {
            if (radix_pass(src, dst, n, shift, sh)) {
                uint2 *t = src; src = dst; dst = t;
            }
        }
        fix_ties(src, n);
    }
    for (int i = threadIdx.x; i < n; i += SORT_THREADS) sorted_gid[off + i] = (int)src[i].y;
}

}  // namespace

cudaError_t sas_launch_tile_sort(uint2 *pairs, uint2 *pairs_alt, const int *tile_offset, int n_buckets, int *sorted_gid,
                                 long long max_isects, cudaStream_t st, int *n_launches) {
    // size classes [lo, hi): shared memory = SortShared + 2*(hi-1) pairs
    static const int bounds[4] = {1, 2049, 6145, 12289};
    static bool attr_set_dev[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    bool &attr_set = attr_set_dev[dev & 63];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tile_sort_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(sizeof(SortShared) + 2 * (size_t)(bounds[3] - 1) * sizeof(uint2)));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    for (int c = 0; c < 3; c++) {
        const size_t smem = sizeof(SortShared) + 2 * (size_t)(bounds[c + 1] - 1) * sizeof(uint2);
        tile_sort_kernel<false><<<n_buckets, SORT_THREADS, smem, st>>>(pairs, pairs_alt, tile_offset, sorted_gid, bounds[c],
                                                                      bounds[c + 1], max_isects);
    }
    tile_sort_kernel<true><<<n_buckets, SORT_THREADS, sizeof(SortShared), st>>>(pairs, pairs_alt, tile_offset, sorted_gid, bounds[3],
                                                                               0x7fffffff, max_isects);
    if (n_launches) *n_launches += 4;
    return cudaGetLastError();
}
