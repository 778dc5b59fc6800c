// mbar_wait_probe.cu — what does it cost a WARP to wait on an mbarrier that has already completed?  (tuning aid)
//
// The gather kernel's movers wait on an mbarrier per 4-env unit.  Phase clocks showed ~900 cycles per wait even when the
// phase had completed long before, whether all 32 lanes executed the try_wait (lane 0 through after ~100 cycles, lane 31
// after ~900) or lane 0 alone followed by __syncwarp.  This probe times the styles in isolation: W warps per block, each
// with its own barrier (count 1); per iteration lane 0 arrives (completing the phase), then the warp waits for that phase.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CHECK(x)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait_loop(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n"
                 ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

template <int STYLE>
__global__ void __launch_bounds__(1024, 1) wait_kernel(const int iters, unsigned long long *out) {
    __shared__ __align__(8) unsigned long long bars[32];
    __shared__ volatile int flag[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar = smem_u32(&bars[warp]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
        flag[warp] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    long long total = 0;
    for (int i = 0; i < iters; ++i) {
        const uint32_t parity = i & 1;
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
        __syncwarp();
        // some unrelated work so that the arrival is long past
        float x = (float)lane;
#pragma unroll 1
        for (int k = 0; k < 64; ++k) x = x * 1.0001f + 0.5f;
        if (x == 12345.0f) flag[warp] = 1;
        __syncwarp();
        int token = i;
        if (STYLE >= 6) {
            // what the gather kernel's movers do: lane 0 works alone for ~1000 cycles (TMA issue), the other lanes wait for it
            // at a shuffle, and right after the shuffle lane 0 alone waits on the mbarrier
            if (lane == 0) {
                float y = 1.0f;
#pragma unroll 1
                for (int k = 0; k < 200; ++k) y = y * 1.0001f + 0.5f;
                if (y == 12345.0f) flag[warp] = 2;
                token = i + 1;
            }
            if (STYLE == 6 || STYLE == 7) token = __shfl_sync(0xFFFFFFFFu, token, 0);
            if (STYLE == 8) { if (lane == 0) flag[warp] = token; __syncwarp(); token = flag[warp]; }
            if (token == -7) break;
        }
        const long long t0 = clock64();
        if (STYLE == 7) __syncwarp();
        if (STYLE == 0) { // every lane waits
            mbar_wait_loop(bar, parity);
        } else if (STYLE == 1 || STYLE >= 6) { // lane 0 waits, __syncwarp
            if (lane == 0) mbar_wait_loop(bar, parity);
        } else if (STYLE == 2) { // every lane polls test_wait
            while (!mbar_test(bar, parity)) {}
        } else if (STYLE == 3) { // lane 0 polls test_wait
            if (lane == 0) while (!mbar_test(bar, parity)) {}
        } else if (STYLE == 4) { // every lane: one try_wait, loop only if it failed
            if (!mbar_try(bar, parity)) mbar_wait_loop(bar, parity);
        } else if (STYLE == 5) { // lane 0: one test, result broadcast by shuffle; fall back to the loop
            int ok = 0;
            if (lane == 0) ok = mbar_test(bar, parity);
            ok = __shfl_sync(0xFFFFFFFFu, ok, 0);
            if (!ok) { if (lane == 0) mbar_wait_loop(bar, parity); }
        }
        __syncwarp();
        const long long t1 = clock64();
        total += t1 - t0;
    }
    if (lane == 0) out[blockIdx.x * 32 + warp] = (unsigned long long)total;
}

template <int STYLE> void run(int warps, const char *what) {
    unsigned long long *d, h[32];
    CHECK(cudaMalloc(&d, 148 * 32 * 8));
    const int iters = 2000;
    wait_kernel<STYLE><<<148, warps * 32>>>(iters, d);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaMemcpy(h, d, 32 * 8, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int w = 0; w < warps; ++w) avg += (double)h[w] / iters;
    printf("style %d  warps %2d  %6.0f cycles per wait (lane 0, wait + __syncwarp)   %s\n", STYLE, warps, avg / warps, what);
    CHECK(cudaFree(d));
}

int main() {
    for (int warps : {1, 12, 20}) {
        run<0>(warps, "every lane try_wait loop");
        run<1>(warps, "lane 0 try_wait loop, __syncwarp");
        run<2>(warps, "every lane test_wait poll");
        run<3>(warps, "lane 0 test_wait poll, __syncwarp");
        run<4>(warps, "every lane one try_wait (+ loop if it failed)");
        run<5>(warps, "lane 0 one test_wait, shuffle broadcast");
        run<6>(warps, "style 1 after 1000 solo cycles of lane 0 + __shfl_sync");
        run<7>(warps, "style 6 with an extra __syncwarp between the shuffle and the wait (inside the timed region)");
        run<8>(warps, "style 1 after 1000 solo cycles of lane 0, broadcast through shared memory + __syncwarp instead of a shuffle");
    }
    return 0;
}
