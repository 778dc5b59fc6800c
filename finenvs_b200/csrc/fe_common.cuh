// fe_common.cuh — pieces shared by the translation units of libfinenvs_b200.so.
#ifndef FE_COMMON_CUH
#define FE_COMMON_CUH
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

// ------------------------------------------------------------------------------------------
// Philox4x32-10 keyed (seed) with counter (env id, step | kind<<63): the redraw RNG
// (replaces torch.randint at :253 / :511; identical on host, see fe_philox)
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint64_t seed, uint64_t env_id, uint64_t step,
                                                       uint32_t kind, uint32_t out[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)env_id, c1 = (uint32_t)(env_id >> 32), c2 = (uint32_t)step,
             c3 = ((uint32_t)(step >> 32) & 0x7FFFFFFFu) | (kind << 31);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

} // namespace
#endif // FE_COMMON_CUH
