// sas_work.cu -- device-side work lists, so the sort and the compositor run as persistent CTAs
// that pull (bucket | tile) ids from an atomic cursor instead of launching one CTA per bucket.
//
// (1) Sort classes: every (row,tile) bucket is appended to the list of its size class (the classes
//     differ in threads and shared memory, sas_sort.cu); empty buckets are dropped.
// (2) Compositor order: tiles are listed heaviest first (32 classes, two per octave of list length;
//     longest-processing-time-first keeps the tail of the kernel short -- measured 1.28x -> 1.01x of
//     the ideal makespan on the pusht batch).  Two kinds of item:
//       * image tile (i,t) whose link bucket is non-empty: composited from its merged list;
//       * view tile (v,t): background-only compositing, done ONCE per view for all the images of
//         that view whose link bucket for tile t is empty (47 % of the pusht batch), and for the
//         4x2-pixel groups of the other tiles that no link Gaussian can reach (tile_dyn_mask, written
//         by the projection kernel) -- their pixels are bit-identical, so they are copied
//         (sas_composite.cu: copy_static_tiles_kernel).
//
// All counts come from the row-tile counts written by the projection kernel; nothing here reads
// back to the host.
#include "sas_common.cuh"

namespace {

__device__ __forceinline__ int sort_class_of(unsigned n) {
    if (n == 0) return -1;
    if (n <= SAS_SORT_CAP0) return 0;
    if (n <= SAS_SORT_CAP1) return 1;
    if (n <= SAS_SORT_CAP2) return 2;
    if (n <= SAS_SORT_CAP3) return 3;
    return 4;
}

// heaviest first: class 0 holds the longest lists
__device__ __forceinline__ int comp_class_of(unsigned n) {
    const int lg = 31 - __clz((int)n);  // n >= 1
    const int half = lg > 0 ? (int)((n >> (lg - 1)) & 1u) : 0;
    const int w = min(31, 2 * lg + half);
    return 31 - w;
}

// pass 1 over [0, RT) row buckets then [RT, RT+IT) image tiles: class histograms + the 'needed' marks
__global__ void __launch_bounds__(256) work_count_kernel(const unsigned *__restrict__ row_tile_count, const int *__restrict__ view_of_image,
                                                         const unsigned *__restrict__ tile_dyn_mask, int C, int I, int T, int tile_w, int W,
                                                         int H, SasWork *work, unsigned char *__restrict__ view_tile_needed) {
    __shared__ unsigned s_sort[SAS_SORT_CLASSES], s_comp[SAS_COMP_CLASSES];
    if (threadIdx.x < SAS_SORT_CLASSES) s_sort[threadIdx.x] = 0;
    if (threadIdx.x < SAS_COMP_CLASSES) s_comp[threadIdx.x] = 0;
    __syncthreads();
    const long long RT = (long long)(C + I) * T, IT = (long long)I * T;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx < RT) {
        const int c = sort_class_of(row_tile_count[idx]);
        if (c >= 0) atomicAdd(&s_sort[c], 1u);
    } else if (idx < RT + IT) {
        const long long j = idx - RT;
        const int i = (int)(j / T), t = (int)(j - (long long)i * T);
        const int v = view_of_image[i];
        const unsigned nd = row_tile_count[(size_t)(C + i) * T + t];
        if (nd > 0) {
            atomicAdd(&s_comp[comp_class_of(nd + row_tile_count[(size_t)v * T + t])], 1u);
            // some visible group is out of reach of every link Gaussian: it is copied from the view tile
            if (~tile_dyn_mask[j] & sas_inside_group_mask(t, tile_w, W, H)) view_tile_needed[(size_t)v * T + t] = 1;
        } else {
            view_tile_needed[(size_t)v * T + t] = 1;  // benign race: every writer stores 1
        }
    }
    __syncthreads();
    if (threadIdx.x < SAS_SORT_CLASSES && s_sort[threadIdx.x]) atomicAdd(&work->sort_count[threadIdx.x], s_sort[threadIdx.x]);
    if (threadIdx.x < SAS_COMP_CLASSES && s_comp[threadIdx.x]) atomicAdd(&work->comp_count[threadIdx.x], s_comp[threadIdx.x]);
}

// pass 2 over the C*T view tiles (needs every 'needed' mark of pass 1)
__global__ void __launch_bounds__(256) work_count_views_kernel(const unsigned *__restrict__ row_tile_count, int C, int T, SasWork *work,
                                                               const unsigned char *__restrict__ view_tile_needed) {
    __shared__ unsigned s_comp[SAS_COMP_CLASSES];
    if (threadIdx.x < SAS_COMP_CLASSES) s_comp[threadIdx.x] = 0;
    __syncthreads();
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx < (long long)C * T && view_tile_needed[idx]) {
        const unsigned n = row_tile_count[idx];
        if (n > 0) atomicAdd(&s_comp[comp_class_of(n)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < SAS_COMP_CLASSES && s_comp[threadIdx.x]) atomicAdd(&work->comp_count[threadIdx.x], s_comp[threadIdx.x]);
}

__global__ void work_prefix_kernel(SasWork *work) {
    if (threadIdx.x == 0) {
        unsigned run = 0;
        for (int c = 0; c < SAS_SORT_CLASSES; c++) { work->sort_start[c] = run; run += work->sort_count[c]; }
        run = 0;
        for (int c = 0; c < SAS_COMP_CLASSES; c++) { work->comp_start[c] = run; run += work->comp_count[c]; }
        work->comp_total = run;
    }
}

// pass 3: same indexing as passes 1 and 2 ([0,RT) buckets, [RT,RT+IT) image tiles, then C*T view tiles)
__global__ void __launch_bounds__(256) work_scatter_kernel(const unsigned *__restrict__ row_tile_count, const int *__restrict__ view_of_image,
                                                           int C, int I, int T, SasWork *work, const unsigned char *__restrict__ view_tile_needed,
                                                           int *__restrict__ sort_list, int *__restrict__ comp_list) {
    const long long RT = (long long)(C + I) * T, IT = (long long)I * T, VT = (long long)C * T;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx < RT) {
        const int c = sort_class_of(row_tile_count[idx]);
        if (c >= 0) sort_list[work->sort_start[c] + atomicAdd(&work->sort_fill[c], 1u)] = (int)idx;
    } else if (idx < RT + IT) {
        const long long j = idx - RT;
        const int i = (int)(j / T), t = (int)(j - (long long)i * T);
        const unsigned nd = row_tile_count[(size_t)(C + i) * T + t];
        if (nd > 0) {
            const int c = comp_class_of(nd + row_tile_count[(size_t)view_of_image[i] * T + t]);
            comp_list[work->comp_start[c] + atomicAdd(&work->comp_fill[c], 1u)] = (int)j;  // image tile: id in [0, IT)
        }
    } else if (idx < RT + IT + VT) {
        const long long k = idx - RT - IT;
        if (view_tile_needed[k]) {
            const unsigned n = row_tile_count[k];
            if (n > 0) {
                const int c = comp_class_of(n);
                comp_list[work->comp_start[c] + atomicAdd(&work->comp_fill[c], 1u)] = (int)(IT + k);  // view tile: id in [IT, IT+VT)
            }
        }
    }
}

// ---- fused single-CTA path -------------------------------------------------------------------------
// Both scans (K2) and all work lists (K2b) in ONE launch of one 1024-thread CTA: for a few thousand
// (row, tile) buckets the eleven small kernels above are pure launch latency (15 % of a single 640x480
// frame), while one CTA streams the counts from L2 in a few microseconds.  Same outputs, same arbitrary
// order inside a class.  Larger problems keep the multi-CTA kernels (sas_scan_lists_fused_ok).
constexpr int FUSED_THREADS = 1024;
constexpr int FUSED_ITEMS = 4;

struct FusedShared {
    unsigned wsum[FUSED_THREADS / 32];
    unsigned carry;
    unsigned sort_count[SAS_SORT_CLASSES], sort_fill[SAS_SORT_CLASSES], sort_start[SAS_SORT_CLASSES];
    unsigned comp_count[SAS_COMP_CLASSES], comp_fill[SAS_COMP_CLASSES], comp_start[SAS_COMP_CLASSES];
    unsigned max_list;
};

// exclusive scan of value(idx), idx in [0, n), written to out[idx]; returns the total (all threads)
template <typename F>
__device__ unsigned long long fused_scan(int n, int *__restrict__ out, FusedShared &sh, F value) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long carry = 0;
    for (int base = 0; base < n; base += FUSED_THREADS * FUSED_ITEMS) {
        const int i0 = base + threadIdx.x * FUSED_ITEMS;
        unsigned v[FUSED_ITEMS], s = 0;
#pragma unroll
        for (int k = 0; k < FUSED_ITEMS; k++) {
            v[k] = (i0 + k < n) ? value(i0 + k) : 0u;
            s += v[k];
        }
        unsigned inc = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) sh.wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {  // scan of the 32 warp totals
            const unsigned w = sh.wsum[lane];
            unsigned winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= d) winc += t;
            }
            sh.wsum[lane] = winc - w;
            if (lane == 31) sh.carry = winc;
        }
        __syncthreads();
        unsigned long long ex = carry + sh.wsum[warp] + (inc - s);
#pragma unroll
        for (int k = 0; k < FUSED_ITEMS; k++) {
            if (i0 + k < n) out[i0 + k] = (int)(ex > 0x7fffffffull ? 0x7fffffffull : ex);
            ex += v[k];
        }
        carry += sh.carry;
        __syncthreads();
    }
    return carry;
}

__global__ void __launch_bounds__(FUSED_THREADS) scan_lists_fused_kernel(const unsigned *__restrict__ row_tile_count, const int *__restrict__ view_of_image,
                                                                         const unsigned *__restrict__ tile_dyn_mask, int C, int I, int T, int tile_w,
                                                                         int W, int H, int *__restrict__ row_tile_offset, int *__restrict__ img_tile_offset,
                                                                         SasCounters *counters, long long cap_src, long long cap_all, SasWork *work,
                                                                         unsigned char *__restrict__ view_tile_needed, int *__restrict__ sort_list,
                                                                         int *__restrict__ comp_list) {
    __shared__ FusedShared sh;
    const int RT = (C + I) * T, IT = I * T, VT = C * T;
    const int tid = threadIdx.x;
    if (tid < SAS_SORT_CLASSES) sh.sort_count[tid] = sh.sort_fill[tid] = 0u;
    if (tid < SAS_COMP_CLASSES) sh.comp_count[tid] = sh.comp_fill[tid] = 0u;
    if (tid == 0) sh.max_list = 0u;
    for (int k = tid; k < VT; k += FUSED_THREADS) view_tile_needed[k] = 0;
    __syncthreads();

    // K2: bucket offsets of the (row, tile) buckets, then of the merged (image, tile) lists = isect_offsets
    const unsigned long long m_src = fused_scan(RT, row_tile_offset, sh, [&](int idx) { return row_tile_count[idx]; });
    const unsigned long long m_all = fused_scan(IT, img_tile_offset, sh, [&](int idx) {
        const int i = idx / T, t = idx - i * T;
        return row_tile_count[(size_t)view_of_image[i] * T + t] + row_tile_count[(size_t)(C + i) * T + t];
    });
    if (tid == 0) {
        row_tile_offset[RT] = (int)(m_src > 0x7fffffffull ? 0x7fffffffull : m_src);
        img_tile_offset[IT] = (int)(m_all > 0x7fffffffull ? 0x7fffffffull : m_all);
        counters->n_isects_src = m_src;
        counters->n_isects = m_all;
        unsigned ov = 0u;
        if (m_src > (unsigned long long)cap_src) ov |= 1u;
        if (m_all > (unsigned long long)cap_all) ov |= 2u;
        if (ov) atomicOr(&counters->overflow, ov);
    }

    // K2b pass 1: class histograms, 'needed' marks, longest merged list
    unsigned mx = 0u;
    for (int idx = tid; idx < RT + IT; idx += FUSED_THREADS) {
        if (idx < RT) {
            const int c = sort_class_of(row_tile_count[idx]);
            if (c >= 0) atomicAdd(&sh.sort_count[c], 1u);
        } else {
            const int j = idx - RT;
            const int i = j / T, t = j - i * T;
            const int v = view_of_image[i];
            const unsigned nd = row_tile_count[(size_t)(C + i) * T + t], nbk = row_tile_count[(size_t)v * T + t];
            mx = max(mx, nd + nbk);
            if (nd > 0) {
                atomicAdd(&sh.comp_count[comp_class_of(nd + nbk)], 1u);
                if (~tile_dyn_mask[j] & sas_inside_group_mask(t, tile_w, W, H)) view_tile_needed[(size_t)v * T + t] = 1;
            } else {
                view_tile_needed[(size_t)v * T + t] = 1;  // benign race: every writer stores 1
            }
        }
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((tid & 31) == 0 && mx) atomicMax(&sh.max_list, mx);
    __syncthreads();  // also orders the global 'needed' marks for this CTA's later reads
    // pass 2: needed view tiles
    for (int k = tid; k < VT; k += FUSED_THREADS)
        if (view_tile_needed[k]) {
            const unsigned n = row_tile_count[k];
            if (n > 0) atomicAdd(&sh.comp_count[comp_class_of(n)], 1u);
        }
    __syncthreads();
    if (tid == 0) {
        unsigned run = 0;
        for (int c = 0; c < SAS_SORT_CLASSES; c++) { sh.sort_start[c] = run; work->sort_start[c] = run; work->sort_count[c] = sh.sort_count[c]; run += sh.sort_count[c]; }
        run = 0;
        for (int c = 0; c < SAS_COMP_CLASSES; c++) { sh.comp_start[c] = run; work->comp_start[c] = run; work->comp_count[c] = sh.comp_count[c]; run += sh.comp_count[c]; }
        work->comp_total = run;
        counters->max_tile_list = sh.max_list;
    }
    __syncthreads();
    // pass 3: scatter
    for (int idx = tid; idx < RT + IT + VT; idx += FUSED_THREADS) {
        if (idx < RT) {
            const int c = sort_class_of(row_tile_count[idx]);
            if (c >= 0) sort_list[sh.sort_start[c] + atomicAdd(&sh.sort_fill[c], 1u)] = idx;
        } else if (idx < RT + IT) {
            const int j = idx - RT;
            const int i = j / T, t = j - i * T;
            const unsigned nd = row_tile_count[(size_t)(C + i) * T + t];
            if (nd > 0) {
                const int c = comp_class_of(nd + row_tile_count[(size_t)view_of_image[i] * T + t]);
                comp_list[sh.comp_start[c] + atomicAdd(&sh.comp_fill[c], 1u)] = j;  // image tile: id in [0, IT)
            }
        } else {
            const int k = idx - RT - IT;
            if (view_tile_needed[k]) {
                const unsigned n = row_tile_count[k];
                if (n > 0) {
                    const int c = comp_class_of(n);
                    comp_list[sh.comp_start[c] + atomicAdd(&sh.comp_fill[c], 1u)] = IT + k;  // view tile: id in [IT, IT+VT)
                }
            }
        }
    }
}

__global__ void snapshot_kernel(const SasCounters *__restrict__ counters, SasSnapRing *ring) {
    if (threadIdx.x == 0) {
        const unsigned long long k = ring->seq;
        ring->slot[k & 3ull] = *counters;
        ring->seq = k + 1ull;
    }
}

}  // namespace

cudaError_t sas_launch_snapshot(const SasCounters *counters, SasSnapRing *ring, cudaStream_t st) {
    snapshot_kernel<<<1, 32, 0, st>>>(counters, ring);
    return cudaGetLastError();
}

bool sas_scan_lists_fused_ok(int C, int I, int T) {
    const long long n = (long long)(C + I) * T + (long long)I * T + (long long)C * T;
    return n <= (1ll << 14) && !(sas_variant_flags() & 0x100);  // measured: 13 vs 41 us at 3.6 k buckets, 37 vs 48 us at 18 k, 329 vs 75 us at 200 k
}

cudaError_t sas_launch_scan_lists_fused(const unsigned int *row_tile_count, const int *view_of_image, const unsigned int *tile_dyn_mask, int C,
                                        int I, int W, int H, int tile_w, int tile_h, int *row_tile_offset, int *img_tile_offset,
                                        SasCounters *counters, long long cap_src, long long cap_all, SasWork *work,
                                        unsigned char *view_tile_needed, int *sort_list, int *comp_list, cudaStream_t st, int *n_launches) {
    scan_lists_fused_kernel<<<1, FUSED_THREADS, 0, st>>>(row_tile_count, view_of_image, tile_dyn_mask, C, I, tile_w * tile_h, tile_w, W, H,
                                                         row_tile_offset, img_tile_offset, counters, cap_src, cap_all, work, view_tile_needed,
                                                         sort_list, comp_list);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

cudaError_t sas_launch_work_lists(const unsigned int *row_tile_count, const int *view_of_image, const unsigned int *tile_dyn_mask, int C, int I,
                                  int W, int H, int tile_w, int tile_h, SasWork *work, unsigned char *view_tile_needed, int *sort_list,
                                  int *comp_list, cudaStream_t st, int *n_launches) {
    const int T = tile_w * tile_h;
    const long long RT = (long long)(C + I) * T, IT = (long long)I * T, VT = (long long)C * T;
    cudaError_t e = cudaMemsetAsync(view_tile_needed, 0, (size_t)VT, st);
    if (e != cudaSuccess) return e;
    work_count_kernel<<<(unsigned)((RT + IT + 255) / 256), 256, 0, st>>>(row_tile_count, view_of_image, tile_dyn_mask, C, I, T, tile_w, W, H, work,
                                                                            view_tile_needed);
    work_count_views_kernel<<<(unsigned)((VT + 255) / 256), 256, 0, st>>>(row_tile_count, C, T, work, view_tile_needed);
    work_prefix_kernel<<<1, 32, 0, st>>>(work);
    work_scatter_kernel<<<(unsigned)((RT + IT + VT + 255) / 256), 256, 0, st>>>(row_tile_count, view_of_image, C, I, T, work, view_tile_needed,
                                                                               sort_list, comp_list);
    if (n_launches) *n_launches += 4;
    return cudaGetLastError();
}
