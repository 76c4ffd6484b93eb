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
// <=512 entries (4 warps), <=2048 (8 warps), <=5632 and <=12288 (16 warps) in shared memory;
// larger buckets run the same passes directly on global memory (ping-pong with a second pair
// buffer) in one CTA of 32 warps.  (Packed 16-bit histograms that let the <=12288 class run 32 warps
// were measured slower: the per-pass fixed costs grow with the warp count.)  Each class is one launch of persistent CTAs that pull bucket ids from the
// class's work list (sas_work.cu), so empty classes cost one near-empty launch.
//
// HBM traffic: 8 B read + 8 B written per bucket entry.
#include "sas_common.cuh"

#include <cstdlib>

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
            // lanes of this round with the same digit: 8 ballots, one per digit bit.  (match.any does this in
            // one instruction but measured ~6500 cycles per round on B200 -- 63 % of the kernel's stall samples.)
            unsigned peers = amask;
#pragma unroll
            for (int bit = 0; bit < 8; bit++) {
                const bool set = (d >> bit) & 1u;
                const unsigned m = __ballot_sync(amask, set);
                peers &= set ? m : ~m;
            }
            const int leader = __ffs(peers) - 1;
            const unsigned rank = __popc(peers & ((1u << lane) - 1u));
            // every lane reads its digit's running position itself (same-address broadcast inside a peer group);
            // a __shfl_sync(peers, ...) from the leader instead executes once per distinct digit -- up to 32
            // serialised convergence points per round, 41 % of the kernel's stall samples
            const unsigned base = sh.hist[warp][d];
            __syncwarp(amask);  // all reads done before the leaders advance the positions
            if (lane == leader) sh.hist[warp][d] = base + __popc(peers);
            dst[base + rank] = kv;
        }
        __syncwarp();
    }
    __syncthreads();
    return true;
}

// Equal-depth runs -> ascending Gaussian id (what a stable sort with ascending-Gaussian emission gives).  Parallel over
// the ELEMENTS of the runs: an element of a run finds the run's ends, counts the ids of the run that are smaller than
// its own (ids are unique inside a bucket) and is written at run start + that rank into the other buffer; elements
// outside runs are copied.  O(run length) per element, all threads busy -- round 1 let the thread at a run's start
// insertion-sort the whole run alone, O(k^2) serial steps on coplanar, fronto-parallel Gaussians.  Returns true if
// the result is in `out` (false: no ties at all, `a` is already final and `out` was not touched).
// canon_of (may be null): the library stores the background Gaussians in a spatially sorted INTERNAL order
// (sas_api.cu: Morton order of the means, so that a CTA's Gaussians fall into neighbouring tiles); the ids in the pairs
// are internal, the tie order the reference defines is by CANONICAL id: canon_of[internal] for ids below n_bg, the id
// itself for link Gaussians.
template <int NT>
__device__ bool fix_ties(const uint2 *a, uint2 *out, int n, const int *__restrict__ canon_of, int n_bg) {
    bool any = false;
    for (int i = threadIdx.x; i < n && !any; i += NT) any = i + 1 < n && a[i].x == a[i + 1].x;
    if (!__syncthreads_or(any)) return false;
    for (int i = threadIdx.x; i < n; i += NT) {
        const uint2 kv = a[i];
        int pos = i;
        if ((i > 0 && a[i - 1].x == kv.x) || (i + 1 < n && a[i + 1].x == kv.x)) {
            int s = i, e = i + 1;
            while (s > 0 && a[s - 1].x == kv.x) s--;
            while (e < n && a[e].x == kv.x) e++;
            int r = 0;
            if (canon_of) {
                const unsigned mine = (int)kv.y < n_bg ? (unsigned)__ldg(canon_of + kv.y) : kv.y;
                for (int j = s; j < e; j++) {
                    const unsigned y = a[j].y;
                    r += ((int)y < n_bg ? (unsigned)__ldg(canon_of + y) : y) < mine ? 1 : 0;
                }
            } else {
                for (int j = s; j < e; j++) r += a[j].y < kv.y ? 1 : 0;
            }
            pos = s + r;
        }
        out[pos] = kv;
    }
    __syncthreads();
    return true;
}

// Persistent CTAs: each pulls the next bucket of its size class from the class's work list
// (sas_work.cu) until the list is exhausted.
template <int NT, bool IN_GLOBAL>
__global__ void __launch_bounds__(NT) tile_sort_kernel(uint2 *pairs, uint2 *pairs_alt, const int *__restrict__ tile_offset,
                                                       const int *__restrict__ sort_list, SasWork *work, int cls, int cap,
                                                       long long max_isects, const int *__restrict__ canon_of, int n_bg) {
    constexpr int SORT_THREADS = NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortShared<NT> &sh = *reinterpret_cast<SortShared<NT> *>(smem_raw);
    __shared__ unsigned s_take;
    const unsigned n_items = work->sort_count[cls];
    const int *list = sort_list + work->sort_start[cls];
    for (;;) {
        __syncthreads();  // previous bucket fully written back; s_take free to overwrite
        if (threadIdx.x == 0) s_take = atomicAdd(&work->sort_take[cls], 1u);
        __syncthreads();
        const unsigned k = s_take;
        if (k >= n_items) break;
        const int b = list[k];
        long long off = tile_offset[b], end = tile_offset[b + 1];
        if (off < 0 || end < off) continue;
        if (end > max_isects) end = max_isects;
        const long long nn = end - off;
        if (nn <= 0 || (!IN_GLOBAL && nn > (long long)cap)) continue;
        const int n = (int)nn;
        uint2 *src, *dst;
        if (IN_GLOBAL) {
            src = pairs + off;
            dst = pairs_alt + off;
        } else {
            src = reinterpret_cast<uint2 *>(smem_raw + sizeof(SortShared<NT>));
            dst = src + cap;
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
            if (fix_ties<NT>(src, dst, n, canon_of, n_bg)) {
                uint2 *t = src; src = dst; dst = t;
            }
        }
        // sorted pairs back in place (the merge kernel needs the depth keys as well as the ids)
        if (src != pairs + off)
            for (int i = threadIdx.x; i < n; i += SORT_THREADS) pairs[off + i] = src[i];
    }
}

template <int NT>
size_t sort_smem(int cap) { return sizeof(SortShared<NT>) + 2 * (size_t)cap * sizeof(uint2); }

int g_sort_sms = 148;
int g_sort_radix = 1;  // SAS_SORT=bin: the bin sort of sas_binsort.cu for the shared-memory classes (measured slower, see there); default: LSD radix

}  // namespace

cudaError_t sas_sort_init(int n_sms) {
    g_sort_sms = n_sms > 0 ? n_sms : 148;
    if (const char *ev = getenv("SAS_SORT")) g_sort_radix = (ev[0] != 'b');
    cudaError_t e = sas_binsort_init();
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tile_sort_kernel<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sort_smem<512>(SAS_SORT_CAP3));
    return e;
}

// The five size classes touch disjoint buckets, so with side streams they run concurrently
// (fork on an event of the caller's stream, join back into it): the step pays for the slowest class
// instead of the sum of five partially filled launches.
cudaError_t sas_launch_tile_sort(uint2 *pairs, uint2 *pairs_alt, const int *row_tile_offset, const int *sort_list, SasWork *work,
                                 long long capacity, const int *canon_of, int n_bg, cudaStream_t st, cudaStream_t *aux, cudaEvent_t *ev,
                                 int *n_launches) {
    // grid = SMs x resident CTAs of the class (threads / shared memory bound)
    const int sms = g_sort_sms;
    const bool fork = aux != nullptr && ev != nullptr && (sas_variant_flags() & 4);
    cudaError_t e;
    if (fork && (e = cudaEventRecord(ev[0], st)) != cudaSuccess) return e;
    // class -> stream: the class with the longest critical path (<= 5632 entries, two CTAs per SM) stays on the caller's stream
    cudaStream_t s_of[SAS_SORT_CLASSES] = {st, st, st, st, st};
    if (fork) {
        const int side[4] = {0, 1, 3, 4};
        for (int k = 0; k < 4; k++) {
            s_of[side[k]] = aux[k];
            if ((e = cudaStreamWaitEvent(aux[k], ev[0], 0)) != cudaSuccess) return e;
        }
    }
    if (!g_sort_radix) {
        if ((e = sas_launch_bin_sort(pairs, row_tile_offset, sort_list, work, capacity, sms, s_of)) != cudaSuccess) return e;
    } else {
    tile_sort_kernel<512, false><<<sms * 2, 512, sort_smem<512>(SAS_SORT_CAP2), s_of[2]>>>(pairs, pairs_alt, row_tile_offset, sort_list, work, 2,
                                                                                         SAS_SORT_CAP2, capacity, canon_of, n_bg);
    tile_sort_kernel<512, false><<<sms, 512, sort_smem<512>(SAS_SORT_CAP3), s_of[3]>>>(pairs, pairs_alt, row_tile_offset, sort_list, work, 3,
                                                                                     SAS_SORT_CAP3, capacity, canon_of, n_bg);
    tile_sort_kernel<256, false><<<sms * 5, 256, sort_smem<256>(SAS_SORT_CAP1), s_of[1]>>>(pairs, pairs_alt, row_tile_offset, sort_list, work, 1,
                                                                                         SAS_SORT_CAP1, capacity, canon_of, n_bg);
    tile_sort_kernel<128, false><<<sms * 16, 128, sort_smem<128>(SAS_SORT_CAP0), s_of[0]>>>(pairs, pairs_alt, row_tile_offset, sort_list, work, 0,
                                                                                          SAS_SORT_CAP0, capacity, canon_of, n_bg);
    }
    // buckets beyond shared memory: 32 warps per CTA (measured on the 6 M-Gaussian 1080p chunk: sort 26.5 -> 19.5 ms)
    tile_sort_kernel<1024, true><<<sms, 1024, sizeof(SortShared<1024>), s_of[4]>>>(pairs, pairs_alt, row_tile_offset, sort_list, work, 4, 0,
                                                                                  capacity, canon_of, n_bg);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (fork)
        for (int k = 0; k < 4; k++) {
            if ((e = cudaEventRecord(ev[1 + k], aux[k])) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(st, ev[1 + k], 0)) != cudaSuccess) return e;
        }
    if (n_launches) *n_launches += 5;
    return cudaGetLastError();
}
