// Index-time tokeniser on the device (the step BEFORE the hot path;
// replaces the per-document regex + Counter loop of BM25.fit, bm25.py:58-67 / extractor.py:15-31, which
// costs ~214 us per document on the host).
//
// Input: the LOWER-CASED documents as one UTF-8 byte blob (lower-casing stays on the host: str.lower() can
// turn two non-ASCII code points, U+0130 and U+212A, into ASCII letters, and only Python knows that table).
// A token is a maximal run of [a-z0-9_] bytes (extractor.py:28 after lower()); every byte >= 0x80 is a
// separator, exactly like the reference's ASCII-only character class.  Term identity = a 64-bit hash of the
// token bytes (FNV-1a + splitmix64 finaliser); the host twin lives in index_build.py.
#include "common.cuh"

namespace {

__device__ __forceinline__ bool is_tok(uint8_t c) {
    return (c >= 'a' && c <= 'z') || (c >= '0' && c <= '9') || c == '_' || (c >= 'A' && c <= 'Z');
}
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void token_flags_kernel(const uint8_t* __restrict__ text, int64_t n, uint8_t* __restrict__ flags) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool t = is_tok(text[i]);
        const bool prev = i > 0 && is_tok(text[i - 1]);
        flags[i] = (t && !prev) ? 1 : 0;
    }
}

__global__ void token_hash_kernel(const uint8_t* __restrict__ text, int64_t n, const int64_t* __restrict__ starts,
                                  int64_t n_tokens, int64_t* __restrict__ hashes) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tokens) return;
    uint64_t h = 0xCBF29CE484222325ull;
    int64_t len = 0;
    for (int64_t p = starts[t]; p < n; ++p, ++len) {
        uint8_t c = text[p];
        if (!is_tok(c)) break;
        if (c >= 'A' && c <= 'Z') c += 32;
        h = (h ^ c) * 0x100000001B3ull;
    }
    hashes[t] = (int64_t)(splitmix64(h ^ ((uint64_t)len << 48)) >> 1);   // 63 bits: non-negative as int64
}

}  // namespace

extern "C" {

int hs_token_flags(const uint8_t* text, int64_t n_bytes, uint8_t* flags, void* stream) {
    HS_REQUIRE(n_bytes >= 0, "hs_token_flags: negative size");
    if (n_bytes == 0) return HS_OK;
    HS_REQUIRE(text != nullptr && flags != nullptr, "hs_token_flags: null pointer");
    int64_t blocks = (n_bytes + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    token_flags_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(text, n_bytes, flags);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

int hs_token_hashes(const uint8_t* text, int64_t n_bytes, const int64_t* starts, int64_t n_tokens, int64_t* hashes,
                    void* stream) {
    HS_REQUIRE(n_bytes >= 0 && n_tokens >= 0, "hs_token_hashes: negative size");
    if (n_tokens == 0) return HS_OK;
    HS_REQUIRE(text != nullptr && starts != nullptr && hashes != nullptr, "hs_token_hashes: null pointer");
    token_hash_kernel<<<(unsigned)((n_tokens + 255) / 256), 256, 0, (cudaStream_t)stream>>>(text, n_bytes, starts,
                                                                                          n_tokens, hashes);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

}  // extern "C"
