// sas_binsort.cu -- K4 (round 2): per-(row,tile) bucket sort by (fp32 depth bits, Gaussian id) as a BIN sort.
//
// Second level of the MSD sort of gsplat's (image | tile | depth) key (first level: the counting partition in
// sas_bin.cu).  A bucket is staged in shared memory and sorted by its full 64-bit (depth bits, id) pair:
//
//   1. min / max of the depth bits over the bucket (warp reductions);
//   2. ONE counting partition into B ~ n bins (B a power of two, at most 4096) of the RANGE-NORMALISED key,
//      bin = (key - min) >> shift with shift chosen so that the occupied range spans the B bins -- monotonic, so
//      bins are in final order; positions inside a bin are handed out by shared-memory atomics (any order);
//   3. every bin (on average ~1-3 entries) is finished by ONE thread with an insertion sort on (depth, id);
//   4. the rare heavy bins (> 32 entries: clustered depths) go back through 1-3 with their own, finer range; a bin whose
//      depths are ALL equal is binned by Gaussian id instead (ids are unique inside a bucket).
//
// The result is the (depth, id)-lexicographic order, which is exactly what the reference produces: a stable sort of
// keys emitted in ascending Gaussian index.  Against round 1's four-pass 8-bit LSD radix sort (sas_sort.cu: per pass a
// per-warp histogram, a digit scan and a ballot multi-split, then tie fixing) this touches every entry three times
// instead of ~nine and has no warp-serial dependency chain: ~40 instead of ~135 lane-instructions per entry.
//
// Buckets beyond shared memory (> SAS_SORT_CAP3 entries) keep the radix kernel of sas_sort.cu (tile_sort_kernel<1024,true>).
//
// MEASURED (round 2b, profiles/r02b_binsort_vs_radix.txt): bit-exact on the whole GPU suite + 100 fuzz cases, but SLOWER
// than the radix sort -- sort + merge stage 0.273 vs 0.203 ms on the pusht batch, 46 vs 18 ms on the 8-camera chunk of the
// 6 M-Gaussian scene.  The estimate above holds for depths spread over the bucket's range; the scenes here put 60 % of
// their Gaussians on planar patches, so inside a 16-pixel tile most depths fall into a few bins, which become "heavy"
// and are refined ONE AFTER THE OTHER by the whole CTA (a dozen barriers + a histogram clear per range), and buckets of
// up to 32 entries are finished by a single thread.  The radix sort does not care about the distribution.  Kept as an
// ablation (SAS_SORT=bin); the default is the radix sort.
#include "sas_common.cuh"

#include <cstdlib>

namespace {

constexpr int BS_SMALL = 32;    // bins up to this many entries are finished by one thread
constexpr int BS_STACK = 96;    // pending heavy ranges

template <int BMAX>
struct BinSortShared {
    unsigned off[BMAX + 1];
    unsigned wsum[32];
    unsigned kmin, kmax;
    int sp;                      // stack pointer (entries in use)
    int st_lo[BS_STACK], st_hi[BS_STACK];
    unsigned char st_mode[BS_STACK];
    unsigned s_take;
};

__device__ __forceinline__ bool bs_less(const uint2 a, const uint2 b) { return a.x < b.x || (a.x == b.x && a.y < b.y); }

__device__ __forceinline__ void bs_insertion(uint2 *a, int s, int e) {
    for (int i = s + 1; i < e; i++) {
        const uint2 v = a[i];
        int j = i - 1;
        while (j >= s && bs_less(v, a[j])) { a[j + 1] = a[j]; j--; }
        a[j + 1] = v;
    }
}

// Sorts A[0, n) by (x, y) ascending into Bf[0, n); A is clobbered.  All NT threads of the CTA call it.
template <int NT, int BMAX>
__device__ void binsort_smem(uint2 *A, uint2 *Bf, int n, BinSortShared<BMAX> &sh) {
    const int tid = threadIdx.x, lane = tid & 31;
    if (n <= BS_SMALL) {  // tiny bucket: one thread
        if (tid == 0) {
            for (int i = 0; i < n; i++) Bf[i] = A[i];
            bs_insertion(Bf, 0, n);
        }
        __syncthreads();
        return;
    }
    if (tid == 0) { sh.sp = 1; sh.st_lo[0] = 0; sh.st_hi[0] = n; sh.st_mode[0] = 0; }
    for (;;) {
        __syncthreads();  // the stack (and the previous range's writes) are visible
        const int sp = sh.sp;
        if (sp == 0) break;
        const int lo = sh.st_lo[sp - 1], hi = sh.st_hi[sp - 1], mode = sh.st_mode[sp - 1];
        const int m = hi - lo;
        __syncthreads();  // everyone has read the top entry
        if (tid == 0) { sh.sp = sp - 1; sh.kmin = 0xffffffffu; sh.kmax = 0u; }
        __syncthreads();
#define BS_KEY(e) (mode ? (e).y : (e).x)
        {   // 1. range of the key
            unsigned mn = 0xffffffffu, mx = 0u;
            for (int i = tid; i < m; i += NT) {
                const unsigned k = BS_KEY(A[lo + i]);
                mn = min(mn, k); mx = max(mx, k);
            }
            mn = __reduce_min_sync(0xffffffffu, mn);
            mx = __reduce_max_sync(0xffffffffu, mx);
            if (lane == 0) { atomicMin(&sh.kmin, mn); atomicMax(&sh.kmax, mx); }
        }
        __syncthreads();
        const unsigned kmin = sh.kmin, range = sh.kmax - kmin;
        if (range == 0u) {
            // every key of the range is equal.  By depth: order by Gaussian id instead (unique).  (By id it cannot happen
            // for m > 1; finishing with an insertion sort keeps even that case correct.)
            __syncthreads();
            if (tid == 0) {
                if (mode == 0) { const int q = sh.sp; sh.st_lo[q] = lo; sh.st_hi[q] = hi; sh.st_mode[q] = 1; sh.sp = q + 1; }
                else { for (int i = lo; i < hi; i++) Bf[i] = A[i]; bs_insertion(Bf, lo, hi); }
            }
            continue;
        }
        // 2. bins: B = smallest power of two >= m (32 .. BMAX); shift so that the occupied key range spans them
        int lgB = 5;
        while ((1 << lgB) < m && (1 << lgB) < BMAX) lgB++;
        const int B = 1 << lgB;
        const int rb = 32 - __clz((int)range);  // bits of the range, 1..32 (range >= 1)
        const int shift = rb > lgB ? rb - lgB : 0;
        for (int b = tid; b <= B; b += NT) sh.off[b] = 0u;
        __syncthreads();
        for (int i = tid; i < m; i += NT) atomicAdd(&sh.off[(BS_KEY(A[lo + i]) - kmin) >> shift], 1u);
        __syncthreads();
        {   // exclusive scan of off[0, B): B / NT consecutive bins per thread (B >= 32; threads beyond B idle)
            const int per = (B + NT - 1) / NT;
            const int b0 = tid * per;
            unsigned s = 0;
            for (int k = 0; k < per; k++)
                if (b0 + k < B) s += sh.off[b0 + k];
            unsigned inc = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            if (lane == 31) sh.wsum[tid >> 5] = inc;
            __syncthreads();
            unsigned pre = 0;
            for (int w = 0; w < (tid >> 5); w++) pre += sh.wsum[w];
            unsigned run = pre + inc - s;
            for (int k = 0; k < per; k++)
                if (b0 + k < B) { const unsigned c = sh.off[b0 + k]; sh.off[b0 + k] = run; run += c; }
        }
        __syncthreads();
        // 3. scatter: slot order inside a bin is arbitrary (step 4 sorts it); afterwards off[b] = END of bin b
        for (int i = tid; i < m; i += NT) {
            const uint2 e = A[lo + i];
            const unsigned pos = atomicAdd(&sh.off[(BS_KEY(e) - kmin) >> shift], 1u);
            Bf[lo + pos] = e;
        }
        __syncthreads();
        // 4. finish the bins; heavy ones go back on the stack (their entries return to A)
        const int sp0 = sh.sp;  // uniform: nobody has pushed yet
        __syncthreads();
        for (int b = tid; b < B; b += NT) {
            const int s = b ? (int)sh.off[b - 1] : 0, e = (int)sh.off[b];
            const int c = e - s;
            if (c <= 1) continue;
            bool pushed = false;
            if (c > BS_SMALL) {
                const int q = atomicAdd(&sh.sp, 1);
                if (q < BS_STACK) { sh.st_lo[q] = lo + s; sh.st_hi[q] = lo + e; sh.st_mode[q] = (unsigned char)mode; pushed = true; }
                else atomicSub(&sh.sp, 1);  // stack full (pathological nesting): finish it here, slowly but correctly
            }
            if (!pushed) bs_insertion(Bf, lo + s, lo + e);
        }
        __syncthreads();
        const int sp1 = min(sh.sp, BS_STACK);
        for (int q = sp0; q < sp1; q++)
            for (int i = sh.st_lo[q] + tid; i < sh.st_hi[q]; i += NT) A[i] = Bf[i];
#undef BS_KEY
    }
}

// Persistent CTAs: each pulls the next bucket of its size class from the class's work list (sas_work.cu).
template <int NT, int BMAX>
__global__ void __launch_bounds__(NT) bin_sort_kernel(uint2 *pairs, const int *__restrict__ tile_offset, const int *__restrict__ sort_list,
                                                      SasWork *work, int cls, int cap, long long max_isects) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BinSortShared<BMAX> &sh = *reinterpret_cast<BinSortShared<BMAX> *>(smem_raw);
    uint2 *A = reinterpret_cast<uint2 *>(smem_raw + ((sizeof(BinSortShared<BMAX>) + 15) & ~(size_t)15));
    uint2 *Bf = A + cap;
    const unsigned n_items = work->sort_count[cls];
    const int *list = sort_list + work->sort_start[cls];
    for (;;) {
        __syncthreads();  // previous bucket fully written back; s_take free to overwrite
        if (threadIdx.x == 0) sh.s_take = atomicAdd(&work->sort_take[cls], 1u);
        __syncthreads();
        const unsigned k = sh.s_take;
        if (k >= n_items) break;
        const int b = list[k];
        long long off = tile_offset[b], end = tile_offset[b + 1];
        if (off < 0 || end < off) continue;
        if (end > max_isects) end = max_isects;
        const long long nn = end - off;
        if (nn <= 1 || nn > (long long)cap) continue;
        const int n = (int)nn;
        for (int i = threadIdx.x; i < n; i += NT) A[i] = pairs[off + i];
        __syncthreads();
        binsort_smem<NT, BMAX>(A, Bf, n, sh);
        // sorted pairs back in place (the merge kernel needs the depth keys as well as the ids)
        for (int i = threadIdx.x; i < n; i += NT) pairs[off + i] = Bf[i];
    }
}

template <int BMAX>
size_t bs_smem(int cap) { return ((sizeof(BinSortShared<BMAX>) + 15) & ~(size_t)15) + 2 * (size_t)cap * sizeof(uint2); }

}  // namespace

cudaError_t sas_binsort_init() {
    cudaError_t e = cudaFuncSetAttribute(bin_sort_kernel<512, 4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bs_smem<4096>(SAS_SORT_CAP3));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(bin_sort_kernel<256, 2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bs_smem<2048>(SAS_SORT_CAP1));
    return e;
}

// The four shared-memory size classes (<= 512, <= 2048, <= 5632, <= 12288 entries) of sas_common.cuh; class c runs on
// stream s_of[c].  Grid = SMs x resident CTAs of the class.
cudaError_t sas_launch_bin_sort(uint2 *pairs, const int *row_tile_offset, const int *sort_list, SasWork *work, long long capacity,
                                int n_sms, cudaStream_t *s_of) {
    bin_sort_kernel<512, 4096><<<n_sms * 2, 512, bs_smem<4096>(SAS_SORT_CAP2), s_of[2]>>>(pairs, row_tile_offset, sort_list, work, 2, SAS_SORT_CAP2,
                                                                                         capacity);
    bin_sort_kernel<512, 4096><<<n_sms, 512, bs_smem<4096>(SAS_SORT_CAP3), s_of[3]>>>(pairs, row_tile_offset, sort_list, work, 3, SAS_SORT_CAP3,
                                                                                     capacity);
    bin_sort_kernel<256, 2048><<<n_sms * 5, 256, bs_smem<2048>(SAS_SORT_CAP1), s_of[1]>>>(pairs, row_tile_offset, sort_list, work, 1, SAS_SORT_CAP1,
                                                                                         capacity);
    bin_sort_kernel<128, 512><<<n_sms * 16, 128, bs_smem<512>(SAS_SORT_CAP0), s_of[0]>>>(pairs, row_tile_offset, sort_list, work, 0, SAS_SORT_CAP0,
                                                                                        capacity);
    return cudaGetLastError();
}
