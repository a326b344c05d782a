// hs_b200 C ABI: handle management, error string, small utility kernels.
#include <stdarg.h>

#include <new>

#include "common.cuh"

static thread_local char g_err[512] = "";

void hs_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

__global__ void u32_max_kernel(const uint32_t* __restrict__ x, int64_t n, uint32_t* out) {
    uint32_t m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = max(m, x[i]);
    for (int s = 16; s >= 1; s >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, s));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

extern "C" {

int hs_abi_version(void) { return HS_ABI_VERSION; }
const char* hs_last_error(void) { return g_err; }

int hs_index_create(int device, int64_t n_docs, int64_t doc_base, hs_index** out) {
    HS_REQUIRE(out != nullptr, "hs_index_create: out is null");
    HS_REQUIRE(n_docs >= 0 && doc_base >= 0 && n_docs + doc_base <= 0xFFFFFFFFll,
               "hs_index_create: doc ids must fit uint32 (n_docs=%lld doc_base=%lld)", (long long)n_docs,
               (long long)doc_base);
    int count = 0;
    HS_CUDA(cudaGetDeviceCount(&count));
    HS_REQUIRE(device >= 0 && device < count, "hs_index_create: no CUDA device %d (have %d)", device, count);
    hs_index* idx = new (std::nothrow) hs_index();
    HS_REQUIRE(idx != nullptr, "hs_index_create: out of host memory");
    idx->device = device;
    idx->num_sms = hs_num_sms(device);
    idx->n_docs = n_docs;
    idx->doc_base = doc_base;
    *out = idx;
    return HS_OK;
}

int hs_index_destroy(hs_index* idx) {
    delete idx;
    return HS_OK;
}

int hs_index_set_dense(hs_index* idx, const float* vectors, int32_t dim, int64_t ld, const float* vnorm) {
    HS_REQUIRE(idx != nullptr, "hs_index_set_dense: idx is null");
    HS_REQUIRE(vectors != nullptr || idx->n_docs == 0, "hs_index_set_dense: vectors is null");
    HS_REQUIRE(vnorm != nullptr || idx->n_docs == 0, "hs_index_set_dense: vnorm is null");
    HS_REQUIRE(dim > 0 && dim <= 1024, "hs_index_set_dense: dim %d not in 1..1024", dim);
    HS_REQUIRE(ld >= dim && (ld % 4) == 0, "hs_index_set_dense: ld %lld must be >= dim and a multiple of 4",
               (long long)ld);
    HS_REQUIRE(((uintptr_t)vectors & 15) == 0, "hs_index_set_dense: vectors must be 16-byte aligned");
    idx->vectors = vectors;
    idx->vnorm = vnorm;
    idx->dim = dim;
    idx->ld = ld;
    hs_gemm_attach_f32(idx);
    return HS_OK;
}

int hs_index_set_csr(hs_index* idx, const int64_t* indptr, const uint32_t* postings, int64_t n_terms,
                     int64_t n_postings) {
    HS_REQUIRE(idx != nullptr, "hs_index_set_csr: idx is null");
    HS_REQUIRE(indptr != nullptr, "hs_index_set_csr: indptr is null");
    HS_REQUIRE(n_terms >= 0 && n_postings >= 0, "hs_index_set_csr: negative size");
    HS_REQUIRE(postings != nullptr || n_postings == 0, "hs_index_set_csr: postings is null");
    HS_REQUIRE(((uintptr_t)postings & 7) == 0, "hs_index_set_csr: postings must be 8-byte aligned");
    idx->indptr = indptr;
    idx->postings = reinterpret_cast<const uint2*>(postings);
    idx->n_terms = n_terms;
    idx->n_postings = n_postings;
    return HS_OK;
}

int hs_index_set_doc_stats(hs_index* idx, const uint32_t* dl, double avgdl, double k1, double b,
                           const double* impact_table, uint32_t max_dl, uint32_t tf_cap, void* stream) {
    HS_REQUIRE(idx != nullptr, "hs_index_set_doc_stats: idx is null");
    HS_REQUIRE(dl != nullptr || idx->n_docs == 0, "hs_index_set_doc_stats: dl is null");
    if (impact_table != nullptr && idx->n_docs > 0) {
        // the scoring kernels index the table by doc length without a bound check: verify the bound once here
        // (index time; ordered on the caller's stream -- the one `dl` was produced on -- then synchronised)
        cudaStream_t st = (cudaStream_t)stream;
        uint32_t* d_max = nullptr;
        uint32_t h_max = 0;
        HS_CUDA(cudaSetDevice(idx->device));
        HS_CUDA(cudaMalloc(&d_max, sizeof(uint32_t)));
        cudaError_t e = cudaMemsetAsync(d_max, 0, sizeof(uint32_t), st);
        if (e == cudaSuccess) {
            u32_max_kernel<<<1024, 256, 0, st>>>(dl, idx->n_docs, d_max);
            e = cudaMemcpyAsync(&h_max, d_max, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(d_max);
        HS_CUDA(e);
        HS_REQUIRE(h_max <= max_dl, "hs_index_set_doc_stats: a doc length (%u) exceeds max_dl (%u)", h_max, max_dl);
        HS_REQUIRE((uint64_t)(max_dl + 1ull) * (tf_cap + 1ull) < (1ull << 32), "hs_index_set_doc_stats: table too large");
    }
    idx->dl = dl;
    idx->avgdl = avgdl;
    idx->k1 = k1;
    idx->b = b;
    idx->impact_table = impact_table;
    idx->max_dl = max_dl;
    idx->tf_cap = tf_cap;
    return HS_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- stats + key utilities
__global__ void stats_reset_kernel(uint32_t* s, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int slot = i & 3;
        s[i] = (slot == HS_STAT_MIN_A || slot == HS_STAT_MIN_B) ? 0xFFFFFFFFu : 0u;
    }
}
__global__ void stats_decode_kernel(const uint32_t* e, float* f, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f[i] = hs_dec_f32(e[i]);
}
__global__ void stats_encode_kernel(const float* f, uint32_t* e, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) e[i] = hs_enc_f32(f[i]);
}
// all-reduce(MAX) form of the stats: (-min_a, max_a, max_b, -min_b), "nothing seen" (NaN) -> -inf
__global__ void stats_to_maxform_kernel(const uint32_t* e, float* f, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int slot = i & 3;
        float v = hs_dec_f32(e[i]);
        if (slot == HS_STAT_MIN_A || slot == HS_STAT_MIN_B) v = -v;
        f[i] = (v != v) ? __int_as_float(0xff800000) : v;
    }
}
__global__ void stats_from_maxform_kernel(const float* f, uint32_t* e, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int slot = i & 3;
        float v = f[i];
        if (slot == HS_STAT_MIN_A || slot == HS_STAT_MIN_B) v = -v;
        e[i] = hs_enc_f32(v);
    }
}
__global__ void fold_minmax_kernel(const float* __restrict__ x, int64_t n, int slot_min, int slot_max,
                                   uint32_t* stats) {
    const int b = blockIdx.y;
    const float* row = x + (int64_t)b * n;
    float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = row[i];
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    for (int m = 16; m >= 1; m >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, m));
        mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, m));
    }
    if ((threadIdx.x & 31) == 0 && mn <= mx) {
        if (slot_min >= 0) atomicMin(&stats[b * 4 + slot_min], hs_enc_f32(mn));
        if (slot_max >= 0) atomicMax(&stats[b * 4 + slot_max], hs_enc_f32(mx));
    }
}
__global__ void keys_unpack_kernel(const uint64_t* keys, int64_t n, float* scores, int64_t* ids) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        uint64_t k = keys[i];
        if (k == 0) {
            scores[i] = 0.0f;
            ids[i] = -1;
        } else {
            scores[i] = hs_dec_f32((uint32_t)(k >> 32));
            ids[i] = (int64_t)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFu));
        }
    }
}

// out[b, :] = 0 except out[b, id - doc_base] = score for every non-empty key (FAISS-style sparse scores)
__global__ void scatter_keys_kernel(const uint64_t* keys, int64_t count, int k, int64_t n, int64_t doc_base, float* out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        const uint64_t key = keys[i];
        if (key != 0) {
            const int64_t id = (int64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFu)) - doc_base;
            if (id >= 0 && id < n) out[(i / k) * n + id] = hs_dec_f32((uint32_t)(key >> 32));
        }
    }
}

extern "C" {

int hs_scatter_keys(const uint64_t* keys, int32_t B, int32_t k, int64_t n_docs, int64_t doc_base, float* out,
                    void* stream) {
    HS_REQUIRE(keys != nullptr && out != nullptr && B > 0 && k > 0 && n_docs >= 0, "hs_scatter_keys: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_docs == 0) return HS_OK;
    HS_CUDA(cudaMemsetAsync(out, 0, (size_t)B * n_docs * sizeof(float), st));
    const int64_t count = (int64_t)B * k;
    scatter_keys_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(keys, count, k, n_docs, doc_base, out);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

int hs_stats_reset(uint32_t* stats_enc, int32_t B, void* stream) {
    HS_REQUIRE(stats_enc != nullptr && B > 0, "hs_stats_reset: bad arguments");
    stats_reset_kernel<<<(4 * B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats_enc, 4 * B);
    HS_LAUNCH_CHECK();
    return HS_OK;
}
int hs_stats_decode(const uint32_t* stats_enc, float* stats, int32_t B, void* stream) {
    HS_REQUIRE(stats_enc != nullptr && stats != nullptr && B > 0, "hs_stats_decode: bad arguments");
    stats_decode_kernel<<<(4 * B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats_enc, stats, 4 * B);
    HS_LAUNCH_CHECK();
    return HS_OK;
}
int hs_stats_encode(const float* stats, uint32_t* stats_enc, int32_t B, void* stream) {
    HS_REQUIRE(stats_enc != nullptr && stats != nullptr && B > 0, "hs_stats_encode: bad arguments");
    stats_encode_kernel<<<(4 * B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, stats_enc, 4 * B);
    HS_LAUNCH_CHECK();
    return HS_OK;
}
int hs_stats_to_maxform(const uint32_t* stats_enc, float* maxform, int32_t B, void* stream) {
    HS_REQUIRE(stats_enc != nullptr && maxform != nullptr && B > 0, "hs_stats_to_maxform: bad arguments");
    stats_to_maxform_kernel<<<(4 * B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats_enc, maxform, 4 * B);
    HS_LAUNCH_CHECK();
    return HS_OK;
}
int hs_stats_from_maxform(const float* maxform, uint32_t* stats_enc, int32_t B, void* stream) {
    HS_REQUIRE(stats_enc != nullptr && maxform != nullptr && B > 0, "hs_stats_from_maxform: bad arguments");
    stats_from_maxform_kernel<<<(4 * B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(maxform, stats_enc, 4 * B);
    HS_LAUNCH_CHECK();
    return HS_OK;
}
int hs_stats_fold_minmax(const float* x, int64_t n, int32_t B, int32_t slot_min, int32_t slot_max,
                         uint32_t* stats_enc, void* stream) {
    HS_REQUIRE(stats_enc != nullptr && B > 0 && B <= 65535 && n >= 0 && slot_min < 4 && slot_max < 4,
               "hs_stats_fold_minmax: bad arguments");
    if (n == 0) return HS_OK;
    HS_REQUIRE(x != nullptr, "hs_stats_fold_minmax: x is null");
    int64_t blocks = (n + 1023) / 1024;
    if (blocks > 296) blocks = 296;
    dim3 grid((unsigned)blocks, (unsigned)B);
    fold_minmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, slot_min, slot_max, stats_enc);
    HS_LAUNCH_CHECK();
    return HS_OK;
}
int hs_keys_unpack(const uint64_t* keys, int64_t count, float* scores, int64_t* doc_ids, void* stream) {
    HS_REQUIRE(keys != nullptr && scores != nullptr && doc_ids != nullptr && count >= 0,
               "hs_keys_unpack: bad arguments");
    if (count == 0) return HS_OK;
    keys_unpack_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(keys, count, scores,
                                                                                          doc_ids);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

}  // extern "C"
