// Index-build primitives on the device (the step before the hot path; replaces what
// BM25.fit does with Python dicts and Counters per document, bm25.py:45-81): a stable LSD radix sort of 64-bit
// (term, doc) keys, an exclusive prefix sum, run-length encoding of the sorted keys ((term, doc) runs -> tf) and
// the per-term posting counts (df) from the boundaries of the sorted term column.  Integer work, HBM-bound:
// coalesced 8-byte loads, shared-memory histograms / ranks, no global atomics on data-dependent addresses.
//
//   sort pass (8-bit digit): histogram per 2048-key tile -> exclusive scan of the bin-major table -> stable scatter
//   (each warp owns 256 consecutive keys of the tile and ranks them 32 at a time with __match_any_sync, so equal
//   digits keep their input order -- the property LSD needs); 24 bytes of traffic per key and pass.
#include "common.cuh"

namespace {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kKeysPerThread = 8;
constexpr int kTile = kSortThreads * kKeysPerThread;      // 2048 keys per block
constexpr int kBins = 256;

// ---------------------------------------------------------------- exclusive scan (int64), 2048 elements per block
__global__ void __launch_bounds__(kSortThreads) scan_block_kernel(const int64_t* __restrict__ in, int64_t* __restrict__ out,
                                                                 int64_t n, int64_t* __restrict__ block_sums) {
    __shared__ int64_t warp_tot[kSortWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)tid * kKeysPerThread;
    int64_t v[kKeysPerThread], sum = 0;
#pragma unroll
    for (int j = 0; j < kKeysPerThread; ++j) {
        v[j] = (base + j < n) ? in[base + j] : 0;
        sum += v[j];
    }
    int64_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int lo = __shfl_up_sync(0xFFFFFFFFu, (int)(incl & 0xFFFFFFFF), d);
        const int hi = __shfl_up_sync(0xFFFFFFFFu, (int)(incl >> 32), d);
        if (lane >= d) incl += ((int64_t)hi << 32) | (uint32_t)lo;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int64_t woff = 0;
    for (int w = 0; w < warp; ++w) woff += warp_tot[w];
    int64_t run = woff + incl - sum;                      // exclusive prefix of this thread's first element
#pragma unroll
    for (int j = 0; j < kKeysPerThread; ++j) {
        if (base + j < n) out[base + j] = run;
        run += v[j];
    }
    if (tid == kSortThreads - 1 && block_sums != nullptr) block_sums[blockIdx.x] = run;
}

__global__ void scan_add_kernel(int64_t* __restrict__ data, int64_t n, const int64_t* __restrict__ block_off) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) data[i] += block_off[i / kTile];
}

size_t scan_ws_elems(int64_t n) {
    size_t tot = 0;
    while (n > kTile) {
        n = (n + kTile - 1) / kTile;
        tot += (size_t)n + 1;
    }
    return tot + 2;
}

// out[i] = sum of in[0..i); in == out allowed.  ws holds the block sums of every level.
int exclusive_scan(const int64_t* in, int64_t* out, int64_t n, int64_t* ws, cudaStream_t st) {
    if (n <= 0) return HS_OK;
    const int64_t nblk = (n + kTile - 1) / kTile;
    if (nblk == 1) {
        scan_block_kernel<<<1, kSortThreads, 0, st>>>(in, out, n, nullptr);
        HS_LAUNCH_CHECK();
        return HS_OK;
    }
    scan_block_kernel<<<(unsigned)nblk, kSortThreads, 0, st>>>(in, out, n, ws);
    HS_LAUNCH_CHECK();
    int rc = exclusive_scan(ws, ws, nblk, ws + nblk + 1, st);
    if (rc != HS_OK) return rc;
    scan_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out, n, ws);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

// ---------------------------------------------------------------- radix sort pass
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                                                 int64_t nblk, int64_t* __restrict__ hist) {
    __shared__ unsigned int h[kBins];
    const int tid = threadIdx.x;
    h[tid] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kTile;
#pragma unroll
    for (int j = 0; j < kKeysPerThread; ++j) {
        const int64_t i = base + j * kSortThreads + tid;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 0xFF], 1u);
    }
    __syncthreads();
    hist[(int64_t)tid * nblk + blockIdx.x] = h[tid];       // bin-major: the scan order is (digit, tile)
}

__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const uint64_t* __restrict__ keys, uint64_t* __restrict__ out,
                                                                    int64_t n, int shift, int64_t nblk,
                                                                    const int64_t* __restrict__ offs) {
    __shared__ unsigned int cnt[kSortWarps][kBins];       // per warp: keys of each digit seen so far / base in the tile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kSortWarps * kBins; i += kSortThreads) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t wbase = (int64_t)blockIdx.x * kTile + (int64_t)warp * (kTile / kSortWarps);
    uint64_t key[kKeysPerThread];
    unsigned int rank[kKeysPerThread];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < kKeysPerThread; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? keys[i] : 0;
        // lanes past the end get private pseudo-digits so that they never pair with a real key
        const unsigned d = ok ? (unsigned)((key[r] >> shift) & 0xFF) : (unsigned)(kBins + lane);
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
        rank[r] = 0;
        if (ok) {
            rank[r] = cnt[warp][d] + __popc(peers & lt);
        }
        __syncwarp();
        if (ok && (peers & lt) == 0) cnt[warp][d] += __popc(peers);      // the lowest lane of each digit group
        __syncwarp();
    }
    __syncthreads();
    {   // exclusive prefix over the warps, per digit: cnt[w][d] becomes the warp's base inside the tile
        unsigned run = 0;
        for (int w = 0; w < kSortWarps; ++w) {
            const unsigned c = cnt[w][tid];
            cnt[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kKeysPerThread; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        if (i < n) {
            const unsigned d = (unsigned)((key[r] >> shift) & 0xFF);
            out[offs[(int64_t)d * nblk + blockIdx.x] + cnt[warp][d] + rank[r]] = key[r];
        }
    }
}

// ---------------------------------------------------------------- run-length encoding of sorted keys
__global__ void rle_flag_kernel(const uint64_t* __restrict__ keys, int64_t n, int64_t* __restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
// pos = exclusive scan of the head flags; heads write their key and start index, the last element the total
__global__ void rle_scatter_kernel(const uint64_t* __restrict__ keys, int64_t n, const int64_t* __restrict__ pos,
                                   uint64_t* __restrict__ uniq, int64_t* __restrict__ start, int64_t* __restrict__ total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool head = i == 0 || keys[i] != keys[i - 1];
    if (head) {
        uniq[pos[i]] = keys[i];
        start[pos[i]] = i;
    }
    if (i == n - 1) {
        const int64_t m = pos[i] + (head ? 1 : 0);
        *total = m;
        start[m] = n;                                       // sentinel: counts[j] = start[j + 1] - start[j]
    }
}
__global__ void rle_counts_kernel(const int64_t* __restrict__ start, const int64_t* __restrict__ total,
                                  int32_t* __restrict__ counts, int64_t cap) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < cap && j < *total) counts[j] = (int32_t)(start[j + 1] - start[j]);
}

// out[i] = first index j in sorted[0..n) with sorted[j] >= q[i]  (term id of a token hash; membership tests)
__global__ void lower_bound_kernel(const int64_t* __restrict__ sorted, int64_t n, const int64_t* __restrict__ q, int64_t m,
                                   int64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int64_t x = q[i];
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted[mid] < x) lo = mid + 1;
        else hi = mid;
    }
    out[i] = lo;
}

// df[t] = number of sorted unique (term << 32 | doc) keys whose term is t, from the boundaries of the term column
__global__ void term_bounds_kernel(const uint64_t* __restrict__ uk, int64_t m, int64_t* __restrict__ first,
                                   int64_t* __restrict__ last) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint64_t t = uk[i] >> 32;
    if (i == 0 || (uk[i - 1] >> 32) != t) first[t] = i;
    if (i == m - 1 || (uk[i + 1] >> 32) != t) last[t] = i + 1;
}
__global__ void term_df_kernel(const int64_t* __restrict__ first, const int64_t* __restrict__ last, int64_t v,
                               int64_t* __restrict__ df) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < v) df[t] = last[t] - first[t];
}

}  // namespace

extern "C" {

size_t hs_scan_workspace_bytes(int64_t n) { return n <= 0 ? 0 : scan_ws_elems(n) * sizeof(int64_t); }

int hs_exclusive_scan_i64(const int64_t* in, int64_t* out, int64_t n, void* workspace, size_t workspace_bytes,
                          void* stream) {
    HS_REQUIRE(n >= 0, "hs_exclusive_scan_i64: negative size");
    if (n == 0) return HS_OK;
    HS_REQUIRE(in != nullptr && out != nullptr, "hs_exclusive_scan_i64: null pointer");
    HS_REQUIRE(n <= kTile || (workspace != nullptr && workspace_bytes >= hs_scan_workspace_bytes(n)),
               "hs_exclusive_scan_i64: workspace too small");
    return exclusive_scan(in, out, n, (int64_t*)workspace, (cudaStream_t)stream);
}

size_t hs_radix_sort_workspace_bytes(int64_t n) {
    if (n <= 0) return 0;
    const int64_t nblk = (n + kTile - 1) / kTile;
    return ((size_t)kBins * nblk + scan_ws_elems((int64_t)kBins * nblk)) * sizeof(int64_t);
}

int hs_radix_sort_u64(uint64_t* keys, uint64_t* tmp, int64_t n, uint32_t byte_mask, void* workspace,
                      size_t workspace_bytes, void* stream) {
    HS_REQUIRE(n >= 0 && byte_mask != 0 && byte_mask <= 0xFF, "hs_radix_sort_u64: bad arguments");
    if (n <= 1) return HS_OK;
    HS_REQUIRE(keys != nullptr && tmp != nullptr && workspace != nullptr &&
                   workspace_bytes >= hs_radix_sort_workspace_bytes(n), "hs_radix_sort_u64: null pointer or workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nblk = (n + kTile - 1) / kTile;
    HS_REQUIRE(nblk <= 0x7FFFFFFF, "hs_radix_sort_u64: too many keys for one call");
    int64_t* hist = (int64_t*)workspace;
    int64_t* sws = hist + (size_t)kBins * nblk;
    uint64_t* src = keys;
    uint64_t* dst = tmp;
    for (int pass = 0; pass < 8; ++pass) {
        if (!((byte_mask >> pass) & 1u)) continue;      // a byte that is equal in every key: the stable pass is the identity
        radix_hist_kernel<<<(unsigned)nblk, kSortThreads, 0, st>>>(src, n, pass * 8, nblk, hist);
        HS_LAUNCH_CHECK();
        int rc = exclusive_scan(hist, hist, (int64_t)kBins * nblk, sws, st);
        if (rc != HS_OK) return rc;
        radix_scatter_kernel<<<(unsigned)nblk, kSortThreads, 0, st>>>(src, dst, n, pass * 8, nblk, hist);
        HS_LAUNCH_CHECK();
        uint64_t* t = src;
        src = dst;
        dst = t;
    }
    if (src != keys) HS_CUDA(cudaMemcpyAsync(keys, src, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    return HS_OK;
}

size_t hs_rle_workspace_bytes(int64_t n) { return n <= 0 ? 0 : ((size_t)n + scan_ws_elems(n)) * sizeof(int64_t); }

int hs_run_length_encode_u64(const uint64_t* sorted_keys, int64_t n, uint64_t* uniq, int64_t* start, int32_t* counts,
                             int64_t* total, void* workspace, size_t workspace_bytes, void* stream) {
    HS_REQUIRE(n >= 0 && total != nullptr, "hs_run_length_encode_u64: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        HS_CUDA(cudaMemsetAsync(total, 0, sizeof(int64_t), st));
        return HS_OK;
    }
    HS_REQUIRE(sorted_keys != nullptr && uniq != nullptr && start != nullptr && counts != nullptr && workspace != nullptr &&
                   workspace_bytes >= hs_rle_workspace_bytes(n), "hs_run_length_encode_u64: null pointer or workspace too small");
    int64_t* pos = (int64_t*)workspace;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    rle_flag_kernel<<<blocks, 256, 0, st>>>(sorted_keys, n, pos);
    HS_LAUNCH_CHECK();
    int rc = exclusive_scan(pos, pos, n, pos + n, st);
    if (rc != HS_OK) return rc;
    rle_scatter_kernel<<<blocks, 256, 0, st>>>(sorted_keys, n, pos, uniq, start, total);
    rle_counts_kernel<<<blocks, 256, 0, st>>>(start, total, counts, n);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

int hs_lower_bound_i64(const int64_t* sorted, int64_t n, const int64_t* queries, int64_t m, int64_t* out, void* stream) {
    HS_REQUIRE(n >= 0 && m >= 0, "hs_lower_bound_i64: negative size");
    if (m == 0) return HS_OK;
    HS_REQUIRE(queries != nullptr && out != nullptr && (sorted != nullptr || n == 0), "hs_lower_bound_i64: null pointer");
    lower_bound_kernel<<<(unsigned)((m + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sorted, n, queries, m, out);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

int hs_term_doc_freqs(const uint64_t* uniq_keys, int64_t m, int64_t n_terms, int64_t* scratch, int64_t* df, void* stream) {
    HS_REQUIRE(m >= 0 && n_terms >= 0, "hs_term_doc_freqs: negative size");
    if (n_terms == 0) return HS_OK;
    HS_REQUIRE(df != nullptr && scratch != nullptr && (uniq_keys != nullptr || m == 0), "hs_term_doc_freqs: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    HS_CUDA(cudaMemsetAsync(scratch, 0, (size_t)2 * n_terms * sizeof(int64_t), st));
    if (m > 0) {
        term_bounds_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(uniq_keys, m, scratch, scratch + n_terms);
        HS_LAUNCH_CHECK();
    }
    term_df_kernel<<<(unsigned)((n_terms + 255) / 256), 256, 0, st>>>(scratch, scratch + n_terms, n_terms, df);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

}  // extern "C"
