// sas_bin.cu -- K2 exclusive scan of the per-(image,tile) counts (= isect_offsets, gsplat
// _isect_offset_encode) with warp-shuffle prefix sums, and K3 emission of (depth bits, gaussian id)
// pairs into each (image,tile) bucket.
//
// Together with sas_sort.cu this is an MSD radix sort of gsplat's 64-bit key
// (image | tile | fp32 depth bits): the (image,tile) digit is resolved here by a counting
// partition (bucket = digit), the 32 depth bits by 8-bit LSD passes in shared memory.  The
// resulting order is bit-identical to a stable sort of the full key with emission in ascending
// Gaussian index, because ties on depth are broken by Gaussian id in sas_sort.cu.
//
// HBM traffic: scan 8 B per (image,tile); emit reads 8 B per (image,gaussian), writes 8 B per
// intersection.
#include "sas_common.cuh"

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

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums(const unsigned *__restrict__ in, int n, unsigned *__restrict__ block_sums,
                                                                SasCounters *counters) {
    __shared__ unsigned wsum[8];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned s = 0, mx = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        unsigned v = (base + k < n) ? in[base + k] : 0u;
        s += v;
        mx = max(mx, v);
    }
    unsigned tot;
    block_exclusive_scan(s, &tot, wsum);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(&counters->max_tile_list, mx);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

// single block: in-place exclusive scan of block_sums[nb]; writes the grand total and the overflow flag
__global__ void __launch_bounds__(SCAN_THREADS) scan_partials(unsigned *block_sums, int nb, SasCounters *counters, long long max_isects) {
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
        counters->n_isects = carry;
        counters->overflow = (carry > (unsigned long long)max_isects) ? 1u : 0u;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_final(const unsigned *__restrict__ in, int n, const unsigned *__restrict__ block_sums,
                                                           int *__restrict__ out, const SasCounters *counters) {
    __shared__ unsigned wsum[8];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
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
        unsigned long long m = counters->n_isects;
        out[n] = (int)(m > 0x7fffffffull ? 0x7fffffffull : m);
    }
}

__global__ void __launch_bounds__(256) emit_kernel(const uint2 *__restrict__ bins, int N, int tile_w, int tile_h,
                                                   const int *__restrict__ tile_offset, unsigned int *__restrict__ tile_cursor,
                                                   uint2 *__restrict__ pairs, long long max_isects) {
    const int g = blockIdx.x * 256 + threadIdx.x;
    const int im = blockIdx.y;
    if (g >= N) return;
    const uint2 bn = __ldg(bins + (size_t)im * N + g);
    if (bn.x == 0u) return;  // culled (a visible Gaussian always has x1 >= 1)
    const int x0 = bn.x & 0xff, y0 = (bn.x >> 8) & 0xff, x1 = (bn.x >> 16) & 0xff, y1 = bn.x >> 24;
    const size_t tbase = (size_t)im * tile_w * tile_h;
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
            const size_t t = tbase + (size_t)y * tile_w + x;
            const long long pos = (long long)__ldg(tile_offset + t) + (long long)atomicAdd(tile_cursor + t, 1u);
            if (pos < max_isects) pairs[pos] = make_uint2(bn.y, (unsigned)g);
        }
}

}  // namespace

cudaError_t sas_launch_scan(const unsigned int *tile_count, int n, int *tile_offset, unsigned int *block_sums,
                            SasCounters *counters, long long max_isects, cudaStream_t st, int *n_launches) {
    const int nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_block_sums<<<nb, SCAN_THREADS, 0, st>>>(tile_count, n, block_sums, counters);
    scan_partials<<<1, SCAN_THREADS, 0, st>>>(block_sums, nb, counters, max_isects);
    scan_final<<<nb, SCAN_THREADS, 0, st>>>(tile_count, n, block_sums, tile_offset, counters);
    if (n_launches) *n_launches += 3;
    return cudaGetLastError();
}

cudaError_t sas_launch_emit(const uint2 *bins, int N, int I, int tile_w, int tile_h, const int *tile_offset,
                            unsigned int *tile_cursor, uint2 *pairs, long long max_isects, cudaStream_t st) {
    dim3 grid((N + 255) / 256, I);
    emit_kernel<<<grid, 256, 0, st>>>(bins, N, tile_w, tile_h, tile_offset, tile_cursor, pairs, max_isects);
    return cudaGetLastError();
}
