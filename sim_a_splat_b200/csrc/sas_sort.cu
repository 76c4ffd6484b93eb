// sas_sort.cu -- K4: per-(image,tile) bucket sort by (fp32 depth bits, Gaussian id).
//
// Second level of the MSD radix sort of gsplat's (image | tile | depth) key (first level: the
// counting partition in sas_bin.cu).  One CTA per bucket: the bucket is staged into shared
// memory, sorted there with four stable 8-bit LSD passes over the depth bits (per-warp
// histograms, a warp-shuffle scan across digits, and a stable multi-split that ranks equal
// digits inside a warp with match.any + popc), equal-depth runs are then ordered by Gaussian id
// (which is what a stable sort with ascending-Gaussian emission produces in the reference), and
// the sorted (depth, id) pairs are written back in place.  Passes whose digit is uniform over the
// bucket (the exponent byte, usually) are skipped.
//
// Buckets are dispatched by size class (threads and shared-memory footprint = occupancy):
// <=512 entries (4 warps), <=2048 (8 warps), <=6144 and <=12288 (16 warps) in shared memory;
// larger buckets run the same passes directly on global memory (ping-pong with a second pair
// buffer) in one CTA.
//
// HBM traffic: 8 B read + 8 B written per bucket entry.
#include "sas_common.cuh"

namespace {

template <int NT>
struct alignas(16) SortShared {
    unsigned hist[NT / 32][256];
    unsigned tot[256];
    unsigned wsum[8];
    int uniform;
};

// One stable LSD pass src -> dst on bits [shift, shift+8) of .x.  Returns false (and leaves dst
// untouched) if every key has the same digit.
template <int NT>
__device__ bool radix_pass(const uint2 *src, uint2 *dst, int n, int shift, SortShared<NT> &sh) {
    constexpr int SORT_THREADS = NT, SORT_WARPS = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < SORT_WARPS * 256; i += SORT_THREADS) (&sh.hist[0][0])[i] = 0u;
    if (tid == 0) sh.uniform = 0;
    __syncthreads();
    // contiguous segment per warp, in rounds of 32 consecutive keys (keeps the pass stable)
    const int seg = (((n + SORT_WARPS - 1) / SORT_WARPS) + 31) & ~31;
    const int lo = min(n, warp * seg), hi = min(n, lo + seg);
    for (int i = lo + lane; i < hi; i += 32) atomicAdd(&sh.hist[warp][(src[i].x >> shift) & 0xffu], 1u);
    __syncthreads();
    // per digit: exclusive prefix over the warps, digit total
    for (int d = tid; d < 256; d += SORT_THREADS) {
        unsigned run = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            unsigned t = sh.hist[w][d];
            sh.hist[w][d] = run;
            run += t;
        }
        sh.tot[d] = run;
        if (run == (unsigned)n) sh.uniform = 1;
    }
    __syncthreads();
    const int uniform = sh.uniform;  // read before anyone can re-enter and reset it
    // warp 0: exclusive scan of the 256 digit totals (8 per lane + a warp-shuffle scan)
    if (warp == 0) {
        unsigned v[8], s = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { v[k] = sh.tot[lane * 8 + k]; s += v[k]; }
        unsigned inc = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        unsigned ex = inc - s;
#pragma unroll
        for (int k = 0; k < 8; k++) { sh.tot[lane * 8 + k] = ex; ex += v[k]; }
    }
    __syncthreads();
    if (uniform) return false;
    for (int d = tid; d < 256; d += SORT_THREADS) {
        const unsigned base = sh.tot[d];
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) sh.hist[w][d] += base;
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
template <int NT>
__device__ void fix_ties(uint2 *a, int n) {
    for (int i = threadIdx.x; i < n; i += NT) {
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

template <int NT, bool IN_GLOBAL>
__global__ void __launch_bounds__(NT) tile_sort_kernel(uint2 *pairs, uint2 *pairs_alt, const int *__restrict__ tile_offset,
                                                       int lo_size, int hi_size, long long max_isects) {
    constexpr int SORT_THREADS = NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortShared<NT> &sh = *reinterpret_cast<SortShared<NT> *>(smem_raw);
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
        src = reinterpret_cast<uint2 *>(smem_raw + sizeof(SortShared<NT>));
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
        fix_ties<NT>(src, n);
    }
    // sorted pairs back in place (the merge kernel needs the depth keys as well as the ids)
    if (src != pairs + off)
        for (int i = threadIdx.x; i < n; i += SORT_THREADS) pairs[off + i] = src[i];
}

}  // namespace

template <int NT>
static size_t sort_smem(int cap) { return sizeof(SortShared<NT>) + 2 * (size_t)cap * sizeof(uint2); }

cudaError_t sas_sort_init() {
    return cudaFuncSetAttribute(tile_sort_kernel<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem<512>(12288));
}

cudaError_t sas_launch_tile_sort(uint2 *pairs, uint2 *pairs_alt, const int *row_tile_offset, int n_buckets, long long capacity,
                                 cudaStream_t st, int *n_launches) {
    // size classes [lo, hi): shared memory = SortShared + 2*(hi-1) pairs
    tile_sort_kernel<128, false><<<n_buckets, 128, sort_smem<128>(512), st>>>(pairs, pairs_alt, row_tile_offset, 1, 513, capacity);
    tile_sort_kernel<256, false><<<n_buckets, 256, sort_smem<256>(2048), st>>>(pairs, pairs_alt, row_tile_offset, 513, 2049, capacity);
    tile_sort_kernel<512, false><<<n_buckets, 512, sort_smem<512>(6144), st>>>(pairs, pairs_alt, row_tile_offset, 2049, 6145, capacity);
    tile_sort_kernel<512, false><<<n_buckets, 512, sort_smem<512>(12288), st>>>(pairs, pairs_alt, row_tile_offset, 6145, 12289, capacity);
    tile_sort_kernel<512, true><<<n_buckets, 512, sizeof(SortShared<512>), st>>>(pairs, pairs_alt, row_tile_offset, 12289, 0x7fffffff,
                                                                               capacity);
    if (n_launches) *n_launches += 5;
    return cudaGetLastError();
}
