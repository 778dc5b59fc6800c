// tma_store_probe.cu — what does the per-SM bulk-store path sustain on its own?  (tuning aid, not part of the library)
//
// The c2 step (1 Mi envs, W = 60, series L2-resident) writes 1.26 GB of observations per launch with ONE 38 400-byte
// cp.async.bulk.global.shared::cta per 32-env tile and at most two of them in flight per SM; with the window loads switched
// off the kernel still takes 0.23 ms where torch's fill_ of the same bytes takes 0.168 ms (profiles/r01_v4_README.txt,
// r01_v6_README.txt).  This probe separates the store path from everything else: persistent blocks (one per SM) write a
// buffer of the same size with bulk stores of S bytes, D stores in flight, straight out of (unwritten) shared memory — no
// bookkeeping, no loads, no shared stores — next to a plain STG.128 fill of the same buffer.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_store_probe tools/tma_store_probe.cu
//   /tmp/tma_store_probe            # prints GB/s per (S, D, threads that wait on the bulk group)
//
// Reading: if (S = 38400, D = 2) lands near 0.23 ms the mover side of fe_pipe_kernel is not what bounds c2 and the fix is
// the store shape (smaller S with larger D, or stores issued from two threads); if it lands near 0.17 ms the bound is the
// interplay with the movers' loads / shared stores.
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

// D stores of S bytes in flight per block; tile t of block b covers bytes [(b + t * gridDim.x) * S, +S)
template <int D>
__global__ void __launch_bounds__(128, 1) bulk_store_kernel(unsigned char *dst, const size_t total, const uint32_t S) {
    extern __shared__ __align__(128) unsigned char smem[];
    const size_t ntiles = total / S;
    if (threadIdx.x != 0) return;
    int slot = 0;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        // the store that last read this slot must have finished reading shared memory (D - 1 newer ones may still run)
        asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(D - 1) : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + t * S),
                     "r"(smem_u32(smem + (size_t)slot * S)), "r"(S)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        slot = slot + 1 == D ? 0 : slot + 1;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(256) stg_fill_kernel(uint4 *dst, const size_t n16) {
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}

template <typename F> float time_ms(F launch, int reps = 20) {
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CHECK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) launch();
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    CHECK(cudaGetLastError());
    float ms = 0;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

template <int D> void run_bulk(unsigned char *buf, size_t total, int sms, uint32_t S) {
    const size_t smem = (size_t)D * S;
    if (smem > 226 * 1024) return;
    CHECK(cudaFuncSetAttribute(bulk_store_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t bytes = total / S * S;
    const float ms = time_ms([&] { bulk_store_kernel<D><<<sms, 128, smem>>>(buf, total, S); });
    printf("bulk store  S = %6u B  D = %d  smem %6zu B   %.4f ms   %7.1f GB/s\n", S, D, smem, ms, bytes / ms / 1e6);
}

int main() {
    int sms = 0;
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t total = (size_t)1048576 * 60 * 5 * 4; // the c2 observation tensor: 1.26 GB
    unsigned char *buf = nullptr;
    CHECK(cudaMalloc(&buf, total));
    printf("%d SMs, %.3f GB per launch\n", sms, total / 1e9);
    {
        const float ms = time_ms([&] { stg_fill_kernel<<<sms * 8, 256>>>((uint4 *)buf, total / 16); });
        printf("STG.128 fill (8 x 256 threads per SM)             %.4f ms   %7.1f GB/s\n", ms, total / ms / 1e6);
    }
    const uint32_t sizes[] = {2400, 4800, 9600, 19200, 38400, 76800};
    for (uint32_t S : sizes) {
        run_bulk<1>(buf, total, sms, S);
        run_bulk<2>(buf, total, sms, S);
        run_bulk<4>(buf, total, sms, S);
        run_bulk<8>(buf, total, sms, S);
    }
    CHECK(cudaFree(buf));
    return 0;
}
