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

// smallest bucket (exclusive) the multi-CTA path takes, from the number of buckets beyond shared memory (n4): with few of
// them the GPU is idle unless each is spread over many CTAs; with many, one CTA per bucket moves fewer bytes
__device__ __forceinline__ long long sas_big_min_bucket(unsigned n4) { return n4 < 296u ? (long long)SAS_SORT_CAP3 : SAS_BIG_MIN_BUCKET; }

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
        // IN_GLOBAL: cap != 0 says the multi-CTA chunk + merge path is on; it takes the buckets of more than
        // BIG_MIN_BUCKET and at most BIG_MAX_BUCKET entries, this one-CTA kernel the rest
        if (IN_GLOBAL && cap != 0 && nn > sas_big_min_bucket(n_items) && nn <= SAS_BIG_MAX_BUCKET) continue;
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

// ---- buckets beyond shared memory: chunks sorted in shared memory by ALL SMs, then multi-CTA merge passes ---------
// (round 2; round 1 sorted such a bucket with ONE CTA ping-ponging through global memory, which leaves most of the GPU
// idle when a small image over a big scene puts every tile in this class.)  A bucket of n > SAS_SORT_CAP3 entries is cut
// into chunks of SAS_SORT_CAP3; big_plan_kernel lists the (bucket, chunk) items; big_chunk_sort_kernel sorts every chunk
// like a shared-memory bucket; big_merge_kernel, launched BIG_MAX_PASSES times, merges runs of 2^(p-1) chunks pairwise
// by merge path, one CTA per output chunk, ping-ponging between `pairs` and `pairs_alt` so that a bucket of P =
// ceil(log2(chunks)) passes ends in `pairs` (its chunk sort writes into the buffer of the right parity).  Buckets of more
// than 2^BIG_MAX_PASSES chunks keep the one-CTA kernel.  Everything is device-driven: the launches exist in every step's
// graph on the big-bucket stream and exit at once when there is nothing to do.
// Measured (profiles/r02b_big_bucket_sort.txt): a 96x96 view of a 2 M-Gaussian scene (lists up to 155 k entries, 36 tiles)
// sorts in 0.43 instead of 0.73 ms; but when thousands of buckets of 12-30 k entries already fill every SM (the 6 M-Gaussian
// 1080p chunk) the one-CTA kernel moves fewer bytes and is faster (17.9 vs 21.8 ms), and the eight extra launches cost
// ~18 us per step when there is nothing to do.  Hence: when MANY buckets are beyond shared memory (>= 2 per SM: one CTA
// each already fills the GPU) only those of more than 4 chunks take this path, when few do, all of them
// (sas_big_min_bucket); and the launches are only enqueued once the renderer has SEEN such lists (sas_api.cu big_hint, from the counters the host
// reads anyway: stats, overflow polls, the snapshots that travel with host outputs).
constexpr int BIG_MAX_PASSES = 6;                                   // up to 64 chunks = 786 432 entries per bucket
constexpr long long BIG_MAX_BUCKET = SAS_BIG_MAX_BUCKET;
constexpr int BIG_TILE = 2048;                                      // outputs per merge sub-tile (8 per thread, 256 threads)

__device__ __forceinline__ int big_passes(int n_chunks) {
    int p = 0;
    while ((1 << p) < n_chunks) p++;
    return p;
}

__global__ void __launch_bounds__(256) big_plan_kernel(const int *__restrict__ tile_offset, const int *__restrict__ sort_list, SasWork *work,
                                                       long long max_isects, int2 *__restrict__ big_items, int items_cap) {
    const unsigned n4 = work->sort_count[4];
    const int *list = sort_list + work->sort_start[4];
    for (unsigned k = blockIdx.x * 256 + threadIdx.x; k < n4; k += gridDim.x * 256) {
        const int b = list[k];
        long long off = tile_offset[b], end = tile_offset[b + 1];
        if (off < 0 || end < off) continue;
        if (end > max_isects) end = max_isects;
        const long long n = end - off;
        if (n <= sas_big_min_bucket(n4) || n > BIG_MAX_BUCKET) continue;
        const int nc = (int)((n + SAS_SORT_CAP3 - 1) / SAS_SORT_CAP3);
        const unsigned base = atomicAdd(&work->big_items, (unsigned)nc);
        for (int c = 0; c < nc; c++)
            if (base + c < (unsigned)items_cap) big_items[base + c] = make_int2(b, c);
    }
}

// bucket b -> (offset, entries, chunks, passes); false if the bucket is not handled by this path
__device__ __forceinline__ bool big_bucket(const int *__restrict__ tile_offset, int b, long long max_isects, unsigned n4, long long &off, int &n,
                                           int &nc, int &P) {
    off = tile_offset[b];
    long long end = tile_offset[b + 1];
    if (off < 0 || end < off) return false;
    if (end > max_isects) end = max_isects;
    const long long nn = end - off;
    if (nn <= sas_big_min_bucket(n4) || nn > BIG_MAX_BUCKET) return false;
    n = (int)nn;
    nc = (n + SAS_SORT_CAP3 - 1) / SAS_SORT_CAP3;
    P = big_passes(nc);
    return true;
}

__global__ void __launch_bounds__(512) big_chunk_sort_kernel(uint2 *pairs, uint2 *pairs_alt, const int *__restrict__ tile_offset,
                                                             const int2 *__restrict__ big_items, int items_cap, SasWork *work,
                                                             long long max_isects, const int *__restrict__ canon_of, int n_bg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortShared<512> &sh = *reinterpret_cast<SortShared<512> *>(smem_raw);
    __shared__ unsigned s_take;
    const unsigned n_items = min(work->big_items, (unsigned)items_cap);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_take = atomicAdd(&work->big_take[0], 1u);
        __syncthreads();
        const unsigned k = s_take;
        if (k >= n_items) break;
        const int2 it = big_items[k];
        long long off; int nb_, nc, P;
        if (!big_bucket(tile_offset, it.x, max_isects, work->sort_count[4], off, nb_, nc, P)) continue;
        const int c0 = it.y * SAS_SORT_CAP3, n = min(SAS_SORT_CAP3, nb_ - c0);
        uint2 *src = reinterpret_cast<uint2 *>(smem_raw + sizeof(SortShared<512>)), *dst = src + SAS_SORT_CAP3;
        for (int i = threadIdx.x; i < n; i += 512) src[i] = pairs[off + c0 + i];
        __syncthreads();
        if (n > 1) {
#pragma unroll 1
            for (int shift = 0; shift < 32; shift += 8) {
                if (radix_pass(src, dst, n, shift, sh)) {
                    uint2 *t = src; src = dst; dst = t;
                }
            }
            if (fix_ties<512>(src, dst, n, canon_of, n_bg)) {
                uint2 *t = src; src = dst; dst = t;
            }
        }
        uint2 *out = ((P & 1) ? pairs_alt : pairs) + off + c0;  // P more buffer flips follow: an even number ends in `pairs`
        for (int i = threadIdx.x; i < n; i += 512) out[i] = src[i];
    }
}

// 64-bit merge key: depth bits, then the CANONICAL Gaussian id (the order a stable sort of ascending emission gives)
__device__ __forceinline__ unsigned long long big_key(const uint2 e, const int *__restrict__ canon_of, int n_bg) {
    const unsigned id = (canon_of && (int)e.y < n_bg) ? (unsigned)__ldg(canon_of + e.y) : e.y;
    return ((unsigned long long)e.x << 32) | id;
}

__global__ void __launch_bounds__(256) big_merge_kernel(uint2 *pairs, uint2 *pairs_alt, const int *__restrict__ tile_offset,
                                                        const int2 *__restrict__ big_items, int items_cap, SasWork *work, int pass,
                                                        long long max_isects, const int *__restrict__ canon_of, int n_bg) {
    __shared__ uint2 sab[BIG_TILE], so[BIG_TILE];  // the two input windows of a sub-tile share sab: wa + wb = its outputs
    __shared__ unsigned s_take;
    __shared__ int s_split[2];
    const unsigned n_items = min(work->big_items, (unsigned)items_cap);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_take = atomicAdd(&work->big_take[pass], 1u);
        __syncthreads();
        const unsigned k = s_take;
        if (k >= n_items) break;
        const int2 it = big_items[k];
        long long off; int n, nc, P;
        if (!big_bucket(tile_offset, it.x, max_isects, work->sort_count[4], off, n, nc, P)) continue;
        if (pass > P) continue;  // this bucket is finished
        // after pass q the bucket lives in `pairs` iff (P - q) is even
        const uint2 *src = (((P - (pass - 1)) & 1) ? pairs_alt : pairs) + off;
        uint2 *dst = (((P - pass) & 1) ? pairs_alt : pairs) + off;
        const long long half = (long long)SAS_SORT_CAP3 << (pass - 1), pairlen = 2 * half;
        const long long o0 = (long long)it.y * SAS_SORT_CAP3;               // first output position of this item in the bucket
        const long long A0 = (o0 / pairlen) * pairlen;                       // start of the pair of runs it belongs to
        const int na = (int)min(half, (long long)n - A0);
        const long long B0 = A0 + half;
        const int nb = (int)max(0ll, min(half, (long long)n - B0));
        const int d_lo = (int)(o0 - A0), d_hi = (int)min((long long)d_lo + SAS_SORT_CAP3, (long long)na + nb);
        const uint2 *A = src + A0, *B = src + B0;
        if (nb == 0) {  // an unpaired run at the end of the bucket: carried over to the other buffer
            for (int d = d_lo + threadIdx.x; d < d_hi; d += 256) dst[A0 + d] = A[d];
            continue;
        }
        for (int d0 = d_lo; d0 < d_hi; d0 += BIG_TILE) {
            const int cnt = min(BIG_TILE, d_hi - d0);
            __syncthreads();
            if (threadIdx.x < 2) {  // merge-path split of the sub-tile's two ends: entries of A among the first d merged entries
                const int d = threadIdx.x == 0 ? d0 : d0 + cnt;
                int lo = max(0, d - nb), hi = min(d, na);
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (big_key(A[mid], canon_of, n_bg) <= big_key(B[d - 1 - mid], canon_of, n_bg)) lo = mid + 1; else hi = mid;
                }
                s_split[threadIdx.x] = lo;
            }
            __syncthreads();
            const int a0 = s_split[0], wa = s_split[1] - a0, b0 = d0 - a0, wb = cnt - wa;
            uint2 *sa = sab, *sb = sab + wa;
            for (int i = threadIdx.x; i < wa; i += 256) sa[i] = A[a0 + i];
            for (int i = threadIdx.x; i < wb; i += 256) sb[i] = B[b0 + i];
            __syncthreads();
            const int t0 = threadIdx.x * 8;
            if (t0 < cnt) {
                int lo = max(0, t0 - wb), hi = min(t0, wa);
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (big_key(sa[mid], canon_of, n_bg) <= big_key(sb[t0 - 1 - mid], canon_of, n_bg)) lo = mid + 1; else hi = mid;
                }
                int ia = lo, ib = t0 - lo;
#pragma unroll
                for (int v = 0; v < 8; v++)
                    if (t0 + v < cnt) {
                        const bool take_a = ib >= wb || (ia < wa && big_key(sa[ia], canon_of, n_bg) <= big_key(sb[ib], canon_of, n_bg));
                        so[t0 + v] = take_a ? sa[ia] : sb[ib];
                        if (take_a) ia++; else ib++;
                    }
            }
            __syncthreads();
            for (int i = threadIdx.x; i < cnt; i += 256) dst[A0 + d0 + i] = so[i];
        }
    }
}

template <int NT>
size_t sort_smem(int cap) { return sizeof(SortShared<NT>) + 2 * (size_t)cap * sizeof(uint2); }

int g_sort_sms = 148;
int g_sort_big = 1;    // SAS_SORT_BIG=0: buckets beyond shared memory on one CTA each (round 1), no chunk + merge path
int g_sort_radix = 1;  // SAS_SORT=bin: the bin sort of sas_binsort.cu for the shared-memory classes (measured slower, see there); default: LSD radix

}  // namespace

cudaError_t sas_sort_init(int n_sms) {
    g_sort_sms = n_sms > 0 ? n_sms : 148;
    if (const char *ev = getenv("SAS_SORT")) g_sort_radix = (ev[0] != 'b');
    if (const char *ev = getenv("SAS_SORT_BIG")) g_sort_big = atoi(ev) != 0;
    cudaError_t e = sas_binsort_init();
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(big_chunk_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem<512>(SAS_SORT_CAP3));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tile_sort_kernel<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sort_smem<512>(SAS_SORT_CAP3));
    return e;
}

// The five size classes touch disjoint buckets, so with side streams they run concurrently
// (fork on an event of the caller's stream, join back into it): the step pays for the slowest class
// instead of the sum of five partially filled launches.
cudaError_t sas_launch_tile_sort(uint2 *pairs, uint2 *pairs_alt, const int *row_tile_offset, const int *sort_list, SasWork *work,
                                 long long capacity, const int *canon_of, int n_bg, int2 *big_items, int big_items_cap, int big_on,
                                 cudaStream_t st, cudaStream_t *aux, cudaEvent_t *ev, int *n_launches) {
    // grid = SMs x resident CTAs of the class (threads / shared memory bound)
    const int sms = g_sort_sms;
    const bool fork = aux != nullptr && ev != nullptr && (sas_variant_flags() & 4);
    cudaError_t e;
    // giant buckets first, while every SM is free: the chunk sort needs a whole SM's shared memory per CTA and would
    // otherwise queue behind the other classes' CTAs
    const bool big = big_on && big_items != nullptr && big_items_cap > 0 && g_sort_big;
    if (big) {
        big_plan_kernel<<<8, 256, 0, st>>>(row_tile_offset, sort_list, work, capacity, big_items, big_items_cap);
        big_chunk_sort_kernel<<<sms, 512, sort_smem<512>(SAS_SORT_CAP3), st>>>(pairs, pairs_alt, row_tile_offset, big_items, big_items_cap, work,
                                                                            capacity, canon_of, n_bg);
    }
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
    if (big) {
        for (int p = 1; p <= BIG_MAX_PASSES; p++)
            big_merge_kernel<<<sms * 2, 256, 0, s_of[4]>>>(pairs, pairs_alt, row_tile_offset, big_items, big_items_cap, work, p, capacity, canon_of,
                                                          n_bg);
        if (n_launches) *n_launches += 2 + BIG_MAX_PASSES;
    }
    // one CTA per bucket through global memory: every bucket beyond shared memory when the multi-CTA path is off, else
    // those it leaves (at most BIG_MIN_BUCKET or more than BIG_MAX_BUCKET entries)
    tile_sort_kernel<1024, true><<<sms, 1024, sizeof(SortShared<1024>), s_of[4]>>>(pairs, pairs_alt, row_tile_offset, sort_list, work, 4,
                                                                                  big ? 1 : 0, capacity, canon_of, n_bg);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (fork)
        for (int k = 0; k < 4; k++) {
            if ((e = cudaEventRecord(ev[1 + k], aux[k])) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(st, ev[1 + k], 0)) != cudaSuccess) return e;
        }
    if (n_launches) *n_launches += 5;
    return cudaGetLastError();
}
