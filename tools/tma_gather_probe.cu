// tma_gather_probe.cu — can the TMA engine do the whole window gather of the c2 step?  (tuning aid, not part of the library)
//
// Round 1 left fe_pipe_kernel at 0.272 ms on c2 against a 0.20 ms write floor; the movers' LDG -> STS path was the wall and
// one cp.async.bulk per env is limited to ~85 cycles per copy per SM.  This probe measures the read side built on
// cp.async.bulk.tensor ... tile::gather4 (sm_100a): ONE instruction fetches FOUR rows of a 2-D tensor by row index.  With a
// tensor map whose row pitch is smaller than its row length (overlapping rows), "row i" is the window that starts at series
// row i, so one gather4 fetches the windows of four envs.
//
// Data movement only (row0 / position feature are inputs), warp-autonomous: every warp owns a ring of S slots, a slot holds
// one UNIT = 4 consecutive envs.  Modes:
//   0  gather4 from a PRE-INTERLEAVED table (5 floats per row: 4 log-returns + a hole for the position feature; P shifted
//      copies so that every window start is 16-byte aligned): the slot is already the output layout, the warp writes the W
//      position features per env, one bulk store per unit.
//   1  as 0 but one 1-D bulk copy per env (4 per unit)                               — the per-copy cost, for comparison
//   2  gather4 from the plain (T, 4) table, smem -> smem 4 -> 5 interleave by the warp, bulk store per unit
//   3  as 2 but one 1-D bulk copy per env
// flags: 1 = no stores, 2 = no loads.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_gather_probe tools/tma_gather_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
// bounded wait: false = gave up (a dropped copy must not hang the box)
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap *map, int c0, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Args {
    const float *plain;          // (T, 4)
    const unsigned char *inter;  // P shifted copies of the (T, 5) table, copy k at byte k * 80 * M
    const int32_t *row0;         // (N)
    const float *pf;             // (N)
    float *obs;                  // (N, W, 5)
    int64_t N;
    int W, S, flags, M, P, lag;
    int *err;
};

// one warp = one pipeline.  smem per warp: [S mbarriers (64 B)] [S in slots] [modes 2/3: 2 out slots]
template <int MODE>
__global__ void __launch_bounds__(1024, 1) probe_kernel(const __grid_constant__ CUtensorMap map, const Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int W = a.W, S = a.S;
    constexpr bool kInter = MODE < 2;
    constexpr bool kGather = (MODE & 1) == 0;
    const uint32_t in_row = (kInter ? 20u : 16u) * W;            // bytes per env in the in slot
    const uint32_t in_pitch = (4 * in_row + 127) & ~127u;
    const uint32_t out_pitch = (80u * W + 127) & ~127u;
    const uint32_t per_warp = 128 + S * in_pitch + (kInter ? 0 : 2 * out_pitch);
    unsigned char *base = smem + (size_t)warp * per_warp;
    const uint32_t bars = smem_u32(base);
    unsigned char *in_ring = base + 128, *out_ring = in_ring + (size_t)S * in_pitch;
    const int64_t nunits = a.N / 4;
    const int64_t first = (int64_t)blockIdx.x * nw + warp, stride = (int64_t)gridDim.x * nw;
    const int n_mine = first < nunits ? (int)((nunits - first + stride - 1) / stride) : 0;
    if (lane == 0) for (int s = 0; s < S; ++s) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    auto load_unit = [&](int i) { // lane 0 only
        const int s = i % S;
        const int64_t env0 = (first + (int64_t)i * stride) * 4;
        const uint32_t dst = smem_u32(in_ring + (size_t)s * in_pitch), bar = bars + 8 * s;
        if (a.flags & 2) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); return; }
        int r[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int row0 = __ldg(a.row0 + env0 + e);
            r[e] = kInter ? (row0 % a.P) * a.M + row0 / a.P : row0;
        }
        mbar_expect_tx(bar, 4 * in_row);
        if (kGather) {
            gather4(dst, &map, 0, r[0], r[1], r[2], r[3], bar);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const void *src = kInter ? (const void *)(a.inter + (size_t)r[e] * 80) : (const void *)(a.plain + (size_t)r[e] * 4);
                bulk_load(dst + e * in_row, src, in_row, bar);
            }
        }
    };
    const int depth = kInter ? S : S; // units loaded ahead
    if (lane == 0) for (int i = 0; i < depth && i < n_mine; ++i) load_unit(i);

    const int rows = 4 * W;
    for (int i = 0; i < n_mine; ++i) {
        const int s = i % S;
        const int64_t unit = first + (int64_t)i * stride, env0 = unit * 4;
        if (!mbar_wait_bounded(bars + 8 * s, (i / S) & 1)) { if (lane == 0) atomicAdd(a.err, 1); return; }
        const float mypf = lane < 4 ? __ldg(a.pf + env0 + lane) : 0.0f;
        unsigned char *slot = in_ring + (size_t)s * in_pitch;
        if (kInter) {
            float *o = reinterpret_cast<float *>(slot);
            for (int b = 0; b < rows; b += 32) { // every lane takes part in the shuffle
                const int r = b + lane;
                const float v = __shfl_sync(0xFFFFFFFFu, mypf, (r < rows ? r : rows - 1) / W);
                if (r < rows) o[5 * r + 4] = v;
            }
        } else {
            const int o_s = i & 1;
            if (i >= 2) { if (lane == 0) bulk_wait_read<1>(); __syncwarp(); }
            const float4 *in = reinterpret_cast<const float4 *>(slot);
            float *o = reinterpret_cast<float *>(out_ring + (size_t)o_s * out_pitch);
            for (int b = 0; b < rows; b += 32) {
                const int r = b + lane;
                const float f = __shfl_sync(0xFFFFFFFFu, mypf, (r < rows ? r : rows - 1) / W);
                if (r < rows) {
                    const float4 v = in[r];
                    o[5 * r] = v.x; o[5 * r + 1] = v.y; o[5 * r + 2] = v.z; o[5 * r + 3] = v.w; o[5 * r + 4] = f;
                }
            }
            slot = reinterpret_cast<unsigned char *>(o);
        }
        fence_async();
        __syncwarp();
        if (lane == 0) {
            if (!(a.flags & 1)) { bulk_store(a.obs + (size_t)env0 * W * 5, smem_u32(slot), 80u * W); bulk_commit(); }
            if (kInter) {
                // slot of unit i - (lag - 1) may be refilled once its store has read it
                const int j = i - (a.lag - 1);
                if (j >= 0 && j + S < n_mine) {
                    if (a.lag == 1) bulk_wait_read<0>(); else if (a.lag == 2) bulk_wait_read<1>(); else bulk_wait_read<2>();
                    load_unit(j + S);
                }
            } else if (i + S < n_mine) {
                load_unit(i + S); // the in slot was consumed by this warp's own loads above
            }
        }
        __syncwarp();
    }
    if (lane == 0) bulk_wait_read<0>();
}

__global__ void check_kernel(const Args a, unsigned long long *bad) {
    const int64_t n = a.N * a.W * 5;
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < n; f += (int64_t)gridDim.x * blockDim.x) {
        const int64_t env = f / (a.W * 5);
        const int rem = (int)(f - env * a.W * 5), j = rem / 5, c = rem % 5;
        const float want = c == 4 ? a.pf[env] : a.plain[((size_t)a.row0[env] + j) * 4 + c];
        if (a.obs[f] != want) atomicAdd(bad, 1ull);
    }
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE>
void run(const CUtensorMap &map, Args a, int sms, int nw, int S, int lag, int flags, unsigned long long *bad_dev) {
    constexpr bool kInter = MODE < 2;
    const uint32_t in_row = (kInter ? 20u : 16u) * a.W;
    const uint32_t in_pitch = (4 * in_row + 127) & ~127u, out_pitch = (80u * a.W + 127) & ~127u;
    const size_t per_warp = 128 + (size_t)S * in_pitch + (kInter ? 0 : 2 * out_pitch);
    const size_t smem = per_warp * nw;
    if (smem > 226 * 1024) return;
    a.S = S; a.lag = lag; a.flags = flags;
    CHECK(cudaFuncSetAttribute(probe_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CHECK(cudaMemset(a.err, 0, sizeof(int)));
    CHECK(cudaMemset(a.obs, 0xFF, (size_t)a.N * a.W * 20));
    auto launch = [&] { probe_kernel<MODE><<<sms, nw * 32, smem>>>(map, a); };
    launch();
    cudaError_t e = cudaDeviceSynchronize();
    int err = 0;
    if (e != cudaSuccess) { printf("mode %d nw %2d S %d lag %d flags %d: %s\n", MODE, nw, S, lag, flags, cudaGetErrorString(e)); exit(2); }
    CHECK(cudaMemcpy(&err, a.err, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) { printf("mode %d nw %2d S %d lag %d flags %d: %d warps timed out on their mbarrier (copy dropped?)\n", MODE, nw, S, lag, flags, err); return; }
    unsigned long long bad = 0;
    if (flags == 0) {
        CHECK(cudaMemset(bad_dev, 0, 8));
        check_kernel<<<sms * 8, 256>>>(a, bad_dev);
        CHECK(cudaMemcpy(&bad, bad_dev, 8, cudaMemcpyDeviceToHost));
    }
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CHECK(cudaEventRecord(e0));
    const int reps = 20;
    for (int i = 0; i < reps; ++i) launch();
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    printf("mode %d  warps %2d  S %d  lag %d  flags %d  smem %6zu B   %.4f ms   out %7.1f GB/s   mismatches %llu\n", MODE, nw, S, lag,
           flags, smem, ms, (double)a.N * a.W * 20 / ms / 1e6, bad);
    fflush(stdout);
}

int main(int argc, char **argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 60;
    const int64_t N = argc > 2 ? atoll(argv[2]) : (1 << 20);
    const int64_t T = argc > 3 ? atoll(argv[3]) : 258048;
    int sms = 0;
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int P = 4;
    const int64_t M = (T + P - 1) / P + 64;
    printf("W %d  N %lld  T %lld  SMs %d\n", W, (long long)N, (long long)T, sms);

    std::vector<float> plain((size_t)T * 4);
    uint64_t x = 88172645463325252ull;
    auto rnd = [&] { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    for (auto &v : plain) v = (float)((double)(rnd() >> 40) / (1 << 24)) - 0.5f;
    std::vector<unsigned char> inter((size_t)P * M * 80 + 4096, 0);
    {
        std::vector<float> t5((size_t)T * 5, -7.0f);
        for (int64_t r = 0; r < T; ++r) memcpy(&t5[r * 5], &plain[r * 4], 16);
        for (int k = 0; k < P; ++k)
            memcpy(inter.data() + (size_t)k * 80 * M, reinterpret_cast<unsigned char *>(t5.data()) + 20 * k, (size_t)(T - k) * 20);
    }
    std::vector<int32_t> row0(N);
    std::vector<float> pf(N);
    for (int64_t i = 0; i < N; ++i) { row0[i] = (int32_t)(rnd() % (uint64_t)(T - W)); pf[i] = (float)(i % 1000) * 0.001f; }

    Args a{};
    float *d_plain; unsigned char *d_inter; int32_t *d_row0; float *d_pf; float *d_obs; int *d_err; unsigned long long *d_bad;
    CHECK(cudaMalloc(&d_plain, plain.size() * 4));
    CHECK(cudaMalloc(&d_inter, inter.size()));
    CHECK(cudaMalloc(&d_row0, N * 4));
    CHECK(cudaMalloc(&d_pf, N * 4));
    CHECK(cudaMalloc(&d_obs, (size_t)N * W * 20));
    CHECK(cudaMalloc(&d_err, 4));
    CHECK(cudaMalloc(&d_bad, 8));
    CHECK(cudaMemcpy(d_plain, plain.data(), plain.size() * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(d_inter, inter.data(), inter.size(), cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(d_row0, row0.data(), N * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(d_pf, pf.data(), N * 4, cudaMemcpyHostToDevice));
    a.plain = d_plain; a.inter = d_inter; a.row0 = d_row0; a.pf = d_pf; a.obs = d_obs; a.N = N; a.W = W; a.M = (int)M; a.P = P; a.err = d_err;

    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
    if (!encode || qres != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap map_inter, map_plain;
    const cuuint32_t estr[2] = {1, 1};
    {
        const cuuint64_t dims[2] = {(cuuint64_t)(20 * W / 8), (cuuint64_t)(P * M)};
        const cuuint64_t strides[1] = {80};
        const cuuint32_t box[2] = {(cuuint32_t)(20 * W / 8), 1};
        CUresult r = encode(&map_inter, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, d_inter, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode interleaved map (inner %d x u64, pitch 80 B, %lld rows): %d\n", 20 * W / 8, (long long)(P * M), (int)r);
        if (r != CUDA_SUCCESS) return 1;
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)(4 * W), (cuuint64_t)(T - W + 1)};
        const cuuint64_t strides[1] = {16};
        const cuuint32_t box[2] = {(cuuint32_t)(4 * W), 1};
        CUresult r = encode(&map_plain, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_plain, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode plain map (inner %d x f32, pitch 16 B): %d\n", 4 * W, (int)r);
        if (r != CUDA_SUCCESS) return 1;
    }

    // correctness + first timings
    run<0>(map_inter, a, sms, 8, 4, 2, 0, d_bad);
    run<1>(map_inter, a, sms, 8, 4, 2, 0, d_bad);
    run<2>(map_plain, a, sms, 8, 4, 2, 0, d_bad);
    run<3>(map_plain, a, sms, 8, 4, 2, 0, d_bad);
    // sweeps
    for (int nw : {4, 8, 12, 16, 24, 32})
        for (int S : {2, 3, 4, 6})
            for (int lag : {1, 2}) {
                if (lag > S) continue;
                run<0>(map_inter, a, sms, nw, S, lag, 0, d_bad);
            }
    for (int nw : {4, 8, 16, 32}) {
        run<0>(map_inter, a, sms, nw, 3, 2, 1, d_bad); // loads only
        run<0>(map_inter, a, sms, nw, 3, 2, 2, d_bad); // stores only
        run<1>(map_inter, a, sms, nw, 3, 2, 0, d_bad);
        run<1>(map_inter, a, sms, nw, 3, 2, 1, d_bad);
    }
    for (int nw : {4, 8, 12, 16, 24})
        for (int S : {2, 3, 4}) {
            run<2>(map_plain, a, sms, nw, S, 1, 0, d_bad);
        }
    for (int nw : {8, 16}) {
        run<2>(map_plain, a, sms, nw, 3, 1, 1, d_bad);
        run<3>(map_plain, a, sms, nw, 3, 1, 0, d_bad);
        run<3>(map_plain, a, sms, nw, 3, 1, 1, d_bad);
    }
    return 0;
}
