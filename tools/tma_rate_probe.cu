// tma_rate_probe.cu — what does ONE SM's TMA engine sustain for window-sized reads out of an L2-resident table?
// (tuning aid, not part of the library; follow-up of tma_gather_probe.cu)
//
// Loads only, no descriptors from memory (row indices come from an LCG in registers), every warp keeps K copies in
// flight in its own ring.  Op types:
//   0  gather4, rows of RB bytes at 16-byte aligned starts (overlapping-row tensor map, pitch 16)
//   1  gather4, rows at 128-byte aligned starts (row index forced to a multiple of 8)
//   2  4 x 1-D bulk copy of RB bytes (16-byte aligned starts)
//   3  1 x 1-D bulk copy of 4*RB contiguous bytes
//   4  2-D tile load: box {RB bytes, 1 row} x 4 separate ops (one per env), 16-byte aligned starts
// Prints ms per launch and bytes per clock per SM for N = 1 Mi windows.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHECK(x)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap *map, int c0, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tile2d(uint32_t dst, const CUtensorMap *map, int c0, int r0, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(r0), "r"(bar) : "memory");
}

// smem per warp: [K mbarriers, 128 B] [K slots of pitch bytes]
__global__ void __launch_bounds__(1024, 1) rate_kernel(const __grid_constant__ CUtensorMap map, const unsigned char *table,
                                                        const int T, const int RB, const int K, const int op, const int units_per_sm,
                                                        int *err, unsigned long long *sink, const int lanes, unsigned long long *clk) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const uint32_t pitch = (4u * RB + 127) & ~127u;
    // every issuing lane owns a ring: [K mbarriers (128 B)] [K slots]
    const int issuer = warp * lanes + lane, nissuers = nw * lanes;
    unsigned char *base = smem + (size_t)issuer * (128 + (size_t)K * pitch);
    const uint32_t bars = smem_u32(base), ring = smem_u32(base + 128);
    const int n_mine = lane < lanes ? (units_per_sm - issuer + nissuers - 1) / nissuers : 0;
    if (lane < lanes) for (int s = 0; s < K; ++s) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (lane >= lanes) return;
    uint32_t x = 0x9E3779B9u * (blockIdx.x * 2048 + issuer + 1);
    long long t_issue = 0, t_wait = 0;
    auto next_row = [&]() { x = x * 1664525u + 1013904223u; return (int)((x >> 8) % (uint32_t)(T - RB / 16 - 8)); };
    auto issue = [&](int i) {
        const int s = i % K;
        const uint32_t dst = ring + s * pitch, bar = bars + 8 * s;
        int r[4];
        for (int e = 0; e < 4; ++e) { r[e] = next_row(); if (op == 1) r[e] &= ~7; }
        mbar_expect_tx(bar, 4u * RB);
        if (op <= 1) gather4(dst, &map, 0, r[0], r[1], r[2], r[3], bar);
        else if (op == 2) { for (int e = 0; e < 4; ++e) bulk_load(dst + e * RB, table + (size_t)r[e] * 16, RB, bar); }
        else if (op == 3) bulk_load(dst, table + (size_t)r[0] * 16, 4u * RB, bar);
        else { for (int e = 0; e < 4; ++e) tile2d(dst + e * ((RB + 127) & ~127), &map, 0, r[e], bar); }
    };
    for (int i = 0; i < K && i < n_mine; ++i) issue(i);
    for (int i = 0; i < n_mine; ++i) {
        const long long t0 = clock64();
        if (!mbar_wait_bounded(bars + 8 * (i % K), (i / K) & 1)) { atomicAdd(err, 1); return; }
        const long long t1 = clock64();
        if (i + K < n_mine) issue(i + K);
        const long long t2 = clock64();
        t_wait += t1 - t0; t_issue += t2 - t1;
    }
    if (blockIdx.x == 0 && issuer == 0) { clk[0] = (unsigned long long)t_wait; clk[1] = (unsigned long long)t_issue; clk[2] = (unsigned long long)n_mine; }
    if (x == 12345u) *sink = x;
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    int sms = 0, clock_khz = 0;
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CHECK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    const int T = 258048 * 5 / 4;   // 20.6 MB table (same footprint as the 4 shifted copies)
    unsigned char *table; int *err; unsigned long long *sink;
    CHECK(cudaMalloc(&table, (size_t)T * 16 + 65536));
    CHECK(cudaMemset(table, 1, (size_t)T * 16 + 65536));
    CHECK(cudaMalloc(&err, 4)); CHECK(cudaMalloc(&sink, 8));
    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
    const int units_per_sm = (1 << 20) / 4 / sms;
    printf("SMs %d  clock %.0f MHz  units (4 windows) per SM %d\n", sms, clock_khz / 1e3, units_per_sm);
    unsigned long long *clk;
    CHECK(cudaMalloc(&clk, 24));
    for (int RB : {1200}) {
        CUtensorMap map;
        const cuuint64_t dims[2] = {(cuuint64_t)(RB / 8), (cuuint64_t)(T - RB / 16)};
        const cuuint64_t strides[1] = {16};
        const cuuint32_t box[2] = {(cuuint32_t)(RB / 8), 1}, estr[2] = {1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, table, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode RB %d failed: %d\n", RB, (int)r); continue; }
        for (int op : {0, 2, 3})
            for (int nw : {1, 2, 4, 8, 16})
                for (int lanes : {1, 2, 4, 8, 32})
                    for (int K : {1, 2, 4}) {
                        const uint32_t pitch = (4u * RB + 127) & ~127u;
                        const size_t smem = (size_t)nw * lanes * (128 + (size_t)K * pitch);
                        if (smem > 226 * 1024) continue;
                        CHECK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        CHECK(cudaMemset(err, 0, 4));
                        auto launch = [&] { rate_kernel<<<sms, nw * 32, smem>>>(map, table, T, RB, K, op, units_per_sm, err, sink, lanes, clk); };
                        launch();
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("RB %d op %d nw %d K %d: %s\n", RB, op, nw, K, cudaGetErrorString(e)); return 2; }
                        int herr = 0;
                        CHECK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
                        if (herr) { printf("RB %4d op %d warps %2d lanes %2d K %d: %d timeouts\n", RB, op, nw, lanes, K, herr); continue; }
                        cudaEvent_t e0, e1;
                        CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
                        launch();
                        CHECK(cudaEventRecord(e0));
                        const int reps = 5;
                        for (int i = 0; i < reps; ++i) launch();
                        CHECK(cudaEventRecord(e1));
                        CHECK(cudaEventSynchronize(e1));
                        float ms = 0;
                        CHECK(cudaEventElapsedTime(&ms, e0, e1));
                        ms /= reps;
                        unsigned long long h[3];
                        CHECK(cudaMemcpy(h, clk, 24, cudaMemcpyDeviceToHost));
                        const double bytes_sm = (double)units_per_sm * 4 * RB, cycles = ms * 1e-3 * clock_khz * 1e3;
                        printf("RB %4d  op %d  warps %2d  lanes %2d  K %d  %.4f ms  %6.2f B/clk/SM  %5.0f clk/op/SM | issuer 0: wait %5.0f  issue %5.0f clk per op (%llu ops)\n",
                               RB, op, nw, lanes, K, ms, bytes_sm / cycles, cycles / units_per_sm, (double)h[0] / h[2], (double)h[1] / h[2], h[2]);
                        fflush(stdout);
                    }
    }
    return 0;
}
