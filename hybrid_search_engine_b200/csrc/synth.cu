// Counter-based synthetic corpus generators; device twin of hybrid_search_engine_b200/synth.py
// (SURVEY.md section 8d).  Integer hashing plus single-rounding float32 operations only, so the
// output is bit-identical to the numpy definition.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t mix(uint64_t seed_key, uint64_t a, uint64_t b) {
    return splitmix64(splitmix64(seed_key + a) + b);
}

__global__ void synth_emb_kernel(float* out, int64_t row0, int64_t n, int dim, int64_t ld, uint64_t seed_key) {
    const float scale = (float)(1.0 / 37837.22723);
    const int64_t total = n * ld;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ld;
        const int c = (int)(i - r * ld);
        float x = 0.0f;
        if (c < dim) {
            const uint64_t h = mix(seed_key, (uint64_t)(row0 + r), (uint64_t)c);
            const int s = (int)(h & 0xFFFF) + (int)((h >> 16) & 0xFFFF) + (int)((h >> 32) & 0xFFFF) + (int)(h >> 48);
            x = __fmul_rn((float)(s - 131070), scale);
        }
        out[i] = x;
    }
}

__global__ void synth_dl_kernel(uint32_t* dl, int64_t doc0, int64_t n, uint64_t seed_key, uint32_t min_len,
                                uint32_t span) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dl[i] = min_len + (uint32_t)(mix(seed_key, (uint64_t)(doc0 + i), 0) % span);
}

// term(h) = min(V-1, #{r : T[r] <= h}) -- upper-bound binary search on the integer threshold table
__device__ __forceinline__ uint32_t zipf_term(uint64_t h, const uint64_t* __restrict__ T, int V) {
    int lo = 0, hi = V;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(T + mid) <= h) lo = mid + 1; else hi = mid;
    }
    return (uint32_t)(lo < V ? lo : V - 1);
}

// one warp per doc; lanes stride over token positions
__global__ void synth_tokens_kernel(uint64_t* keys, const int64_t* __restrict__ tok_off,
                                    const uint32_t* __restrict__ dl, int64_t doc0, int64_t n, uint64_t seed_key,
                                    const uint64_t* __restrict__ T, int V) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n) return;
    const int64_t off = tok_off[w];
    const uint32_t len = dl[w];
    const uint64_t dkey = splitmix64(seed_key + (uint64_t)(doc0 + w));
    for (uint32_t j = lane; j < len; j += 32) {
        const uint64_t h = splitmix64(dkey + j);
        keys[off + j] = ((uint64_t)zipf_term(h, T, V) << 32) | (uint64_t)w;
    }
}

}  // namespace

extern "C" {

int hs_synth_embeddings(float* out, int64_t row0, int64_t n, int32_t dim, int64_t ld, uint64_t seed_key,
                        void* stream) {
    HS_REQUIRE(n >= 0 && dim > 0 && ld >= dim, "hs_synth_embeddings: bad shape");
    if (n == 0) return HS_OK;
    HS_REQUIRE(out != nullptr, "hs_synth_embeddings: out is null");
    int64_t blocks = (n * ld + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    synth_emb_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(out, row0, n, dim, ld, seed_key);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

int hs_synth_doc_lengths(uint32_t* dl, int64_t doc0, int64_t n, uint64_t seed_key, uint32_t min_len,
                         uint32_t span, void* stream) {
    HS_REQUIRE(n >= 0 && span > 0, "hs_synth_doc_lengths: bad arguments");
    if (n == 0) return HS_OK;
    HS_REQUIRE(dl != nullptr, "hs_synth_doc_lengths: dl is null");
    synth_dl_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dl, doc0, n, seed_key, min_len,
                                                                                  span);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

int hs_synth_token_keys(uint64_t* keys, const int64_t* tok_off, const uint32_t* dl, int64_t doc0, int64_t n,
                        uint64_t seed_key, const uint64_t* thresholds, int32_t vocab, void* stream) {
    HS_REQUIRE(n >= 0 && vocab > 0, "hs_synth_token_keys: bad arguments");
    if (n == 0) return HS_OK;
    HS_REQUIRE(keys != nullptr && tok_off != nullptr && dl != nullptr && thresholds != nullptr,
               "hs_synth_token_keys: null pointer");
    const int64_t blocks = (n * 32 + 255) / 256;
    synth_tokens_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(keys, tok_off, dl, doc0, n, seed_key,
                                                                           thresholds, vocab);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

}  // extern "C"
