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
    int64_t T;
    unsigned long long *clk;
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


// ---------------------------------------------------------------------------------------------------------------------
// v2.  Arch A2: as mode 0 (gather4 from the pre-interleaved table, warp-autonomous) but a unit is U = 4 * G envs
//      (G gather4 loads, ONE store) and the descriptors (row index, position feature) of the warp's next 32 units are
//      fetched in one go (the real kernel gets them from its bookkeeper warps through shared memory).
//      Arch B: block-cooperative 32-env tiles: 1 producer warp (8 gather4 per tile on one mbarrier), C consumer warps
//      (write the position features, fence, arrive), 1 store warp (one 20*W*32-byte bulk store per tile; frees stages).
// ---------------------------------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(1024, 1) probe_a2_kernel(const __grid_constant__ CUtensorMap map, const Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int W = a.W, S = a.S;
    constexpr int U = 4 * G;
    const uint32_t row_b = 20u * W, grp_b = 4 * row_b, grp_pitch = (grp_b + 127) & ~127u, unit_b = U * row_b, pitch = G * grp_pitch;
    const uint32_t per_warp = 128 + S * pitch;
    unsigned char *base = smem + (size_t)warp * per_warp;
    const uint32_t bars = smem_u32(base);
    unsigned char *ring = base + 128;
    const int64_t nunits = a.N / U;
    const int64_t first = (int64_t)blockIdx.x * nw + warp, stride = (int64_t)gridDim.x * nw;
    const int n_mine = first < nunits ? (int)((nunits - first + stride - 1) / stride) : 0;
    if (lane == 0) for (int s = 0; s < S; ++s) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    // descriptor batches: lane l holds unit (batch * 32 + l)'s row indices / features in registers
    int r_reg[U]; float pf_reg[U];
    int batch_loaded = -1;
    auto fetch_batch = [&](int b) {
        const int i = b * 32 + lane;
        if (i < n_mine) {
            const int64_t env0 = (first + (int64_t)i * stride) * U;
#pragma unroll
            for (int e = 0; e < U; ++e) {
                const int row0 = __ldg(a.row0 + env0 + e);
                r_reg[e] = (row0 % a.P) * a.M + row0 / a.P;
                pf_reg[e] = __ldg(a.pf + env0 + e);
            }
        }
        batch_loaded = b;
    };
    // the issue side runs `S` units ahead of the consume side and may sit in the next batch: keep two register sets
    int r_nxt[U]; float pf_nxt[U];
    auto load_unit = [&](int i, const int (&rr)[U]) { // all lanes call; lane 0 issues
        const int s = i % S, src_lane = i & 31;
        const uint32_t dst = smem_u32(ring + (size_t)s * pitch), bar = bars + 8 * s;
        int r[U];
#pragma unroll
        for (int e = 0; e < U; ++e) r[e] = __shfl_sync(0xFFFFFFFFu, rr[e], src_lane);
        if (lane == 0) {
            if (a.flags & 2) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); return; }
            mbar_expect_tx(bar, unit_b);
#pragma unroll
            for (int g = 0; g < G; ++g) gather4(dst + g * grp_pitch, &map, 0, r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3], bar);
        }
    };
    fetch_batch(0);
#pragma unroll
    for (int e = 0; e < U; ++e) { r_nxt[e] = r_reg[e]; pf_nxt[e] = pf_reg[e]; }
    int nxt_batch = 0;
    auto issue = [&](int i) { // unit i may be in batch_loaded or in the next one
        const int b = i >> 5;
        if (b != nxt_batch) { // fetch the next batch into the *_nxt registers
            const int ii = b * 32 + lane;
            if (ii < n_mine) {
                const int64_t env0 = (first + (int64_t)ii * stride) * U;
#pragma unroll
                for (int e = 0; e < U; ++e) {
                    const int row0 = __ldg(a.row0 + env0 + e);
                    r_nxt[e] = (row0 % a.P) * a.M + row0 / a.P;
                    pf_nxt[e] = __ldg(a.pf + env0 + e);
                }
            }
            nxt_batch = b;
        }
        load_unit(i, r_nxt);
    };
    for (int i = 0; i < S && i < n_mine; ++i) issue(i);
    const int rows = U * W;
    for (int i = 0; i < n_mine; ++i) {
        const int s = i % S;
        if ((i >> 5) != batch_loaded) { // consume side enters a new batch: it is the one the issue side already holds
#pragma unroll
            for (int e = 0; e < U; ++e) { r_reg[e] = r_nxt[e]; pf_reg[e] = pf_nxt[e]; }
            batch_loaded = i >> 5;
            if (nxt_batch != batch_loaded) { fetch_batch(batch_loaded); }
        }
        const int64_t env0 = (first + (int64_t)i * stride) * U;
        if (!mbar_wait_bounded(bars + 8 * s, (i / S) & 1)) { if (lane == 0) atomicAdd(a.err, 1); return; }
        float pf_u[U];
#pragma unroll
        for (int e = 0; e < U; ++e) pf_u[e] = __shfl_sync(0xFFFFFFFFu, pf_reg[e], i & 31);
        float *o = reinterpret_cast<float *>(ring + (size_t)s * pitch);
        for (int b = 0; b < rows; b += 32) {
            const int r = b + lane;
            if (r < rows) {
                const int e = r / W, g = e >> 2;
                float v = pf_u[0];
#pragma unroll
                for (int q = 1; q < U; ++q) v = e == q ? pf_u[q] : v;
                reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(o) + g * grp_pitch)[5 * (r - g * 4 * W) + 4] = v;
            }
        }
        fence_async();
        __syncwarp();
        if (lane == 0 && !(a.flags & 1)) {
#pragma unroll
            for (int g = 0; g < G; ++g) bulk_store(a.obs + ((size_t)env0 + 4 * g) * W * 5, smem_u32(o) + g * grp_pitch, grp_b);
            bulk_commit();
        }
        const int j = i - (a.lag - 1);
        if (j >= 0 && j + S < n_mine) {
            if (lane == 0) { if (a.lag == 1) bulk_wait_read<0>(); else if (a.lag == 2) bulk_wait_read<1>(); else bulk_wait_read<2>(); }
            __syncwarp();
            issue(j + S);
        }
    }
    if (lane == 0) bulk_wait_read<0>();
}

// Arch B.  smem: [bars 256 B][pf S x 32 floats][S stages of 32 envs]
__global__ void __launch_bounds__(1024, 1) probe_b_kernel(const __grid_constant__ CUtensorMap map, const Args a, const int C) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = a.W, S = a.S;
    const uint32_t row_b = 20u * W, grp_b = 4 * row_b, grp_pitch = (grp_b + 127) & ~127u, tile_b = 32 * row_b, tile_pitch = 8 * grp_pitch;
    const uint32_t bars = smem_u32(smem);
    auto full = [&](int s) { return bars + 8u * s; };
    auto ready = [&](int s) { return bars + 8u * (8 + s); };
    auto empty = [&](int s) { return bars + 8u * (16 + s); };
    float *pf_s = reinterpret_cast<float *>(smem + 256);
    unsigned char *ring = smem + 256 + 8 * 32 * 4;
    const int64_t ntiles_all = a.N / 32;
    const int ntiles = (int)((ntiles_all - blockIdx.x + gridDim.x - 1) / gridDim.x);
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full(s), 1); mbar_init(ready(s), C); mbar_init(empty(s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto tile_env0 = [&](int t) { return ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * 32; };
    if (warp == 0) {
        // ---------------------------------------------------------------- producer
        int r_next = 0; float pf_next = 0.f;
        if (ntiles > 0) { const int row0 = __ldg(a.row0 + tile_env0(0) + lane); r_next = (row0 % a.P) * a.M + row0 / a.P; pf_next = __ldg(a.pf + tile_env0(0) + lane); }
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % S;
            const int r_cur = r_next; const float pf_cur = pf_next;
            if (t + 1 < ntiles) { const int row0 = __ldg(a.row0 + tile_env0(t + 1) + lane); r_next = (row0 % a.P) * a.M + row0 / a.P; pf_next = __ldg(a.pf + tile_env0(t + 1) + lane); }
            if (t >= S && !mbar_wait_bounded(empty(s), ((t / S) - 1) & 1)) { if (lane == 0) atomicAdd(a.err, 1); return; }
            pf_s[s * 32 + lane] = pf_cur;
            const int r0 = __shfl_sync(0xFFFFFFFFu, r_cur, (lane & 7) * 4), r1 = __shfl_sync(0xFFFFFFFFu, r_cur, (lane & 7) * 4 + 1),
                      r2 = __shfl_sync(0xFFFFFFFFu, r_cur, (lane & 7) * 4 + 2), r3 = __shfl_sync(0xFFFFFFFFu, r_cur, (lane & 7) * 4 + 3);
            __syncwarp();
            if (a.flags & 2) { if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full(s)) : "memory"); continue; }
            if (lane == 0) mbar_expect_tx(full(s), tile_b);
            __syncwarp();
            if (lane < 8) gather4(smem_u32(ring + (size_t)s * tile_pitch) + lane * grp_pitch, &map, 0, r0, r1, r2, r3, full(s));
        }
    } else if (warp <= C) {
        // ---------------------------------------------------------------- consumers: position features of their rows
        const int c = warp - 1, rows = 32 * W;
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % S;
            if (!mbar_wait_bounded(full(s), (t / S) & 1)) { if (lane == 0) atomicAdd(a.err, 1); return; }
            unsigned char *o = ring + (size_t)s * tile_pitch;
            for (int r = c * 32 + lane; r < rows; r += C * 32) {
                const int e = r / W, g = e >> 2;
                reinterpret_cast<float *>(o + g * grp_pitch)[5 * (r - g * 4 * W) + 4] = pf_s[s * 32 + e];
            }
            fence_async();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ready(s)) : "memory");
        }
    } else if (warp == C + 1) {
        // ---------------------------------------------------------------- store warp
        if (lane != 0) return;
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % S;
            if (!mbar_wait_bounded(ready(s), (t / S) & 1)) { atomicAdd(a.err, 1); return; }
            if (!(a.flags & 1)) {
                for (int g = 0; g < 8; ++g)
                    bulk_store(a.obs + ((size_t)tile_env0(t) + 4 * g) * W * 5, smem_u32(ring + (size_t)s * tile_pitch) + g * grp_pitch, grp_b);
                bulk_commit();
            }
            if (t >= 1) { // store t-1 has been read out: its stage is free
                bulk_wait_read<1>();
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty((t - 1) % S)) : "memory");
            }
        }
        bulk_wait_read<0>();
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// v3.  Arch A3: warp-autonomous, a unit is G groups of 4 envs; lane g < G issues ITS group's gather4 and ITS group's bulk
// store (the rate probe showed ~470 cycles of issue latency per TMA op per thread, overlapping across lanes / warps).
// Row indices and features are hashes of the env id (no descriptor traffic; the real kernel gets them from shared memory).
// Phase clocks of block 0 / warp 0 / lane 0 go to a.clk[0..6].
// ---------------------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t hash_row(uint32_t env, uint32_t span) {
    uint32_t x = env * 0x9E3779B9u + 0x7F4A7C15u; x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12; x *= 0x297A2D39u; x ^= x >> 15;
    return x % span;
}
template <int G>
__global__ void __launch_bounds__(1024, 1) probe_a3_kernel(const __grid_constant__ CUtensorMap map, const Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int W = a.W, S = a.S;
    constexpr int U = 4 * G;
    const uint32_t row_b = 20u * W, grp_b = 4 * row_b, grp_pitch = (grp_b + 127) & ~127u, pitch = G * grp_pitch;
    unsigned char *base = smem + (size_t)warp * (128 + S * pitch);
    const uint32_t bars = smem_u32(base);
    unsigned char *ring = base + 128;
    const int64_t nunits = a.N / U;
    const int64_t first = (int64_t)blockIdx.x * nw + warp, stride = (int64_t)gridDim.x * nw;
    const int n_mine = first < nunits ? (int)((nunits - first + stride - 1) / stride) : 0;
    if (lane == 0) for (int s = 0; s < S; ++s) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const uint32_t span = (uint32_t)(a.T - W);
    auto issue = [&](int i) { // converged call; lanes < G issue
        const int s = i % S;
        const uint32_t dst = smem_u32(ring + (size_t)s * pitch), bar = bars + 8 * s;
        if (a.flags & 2) { if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); return; }
        if (lane == 0) mbar_expect_tx(bar, U * row_b);
        __syncwarp();
        if (lane < G) {
            const uint32_t env = (uint32_t)((first + (int64_t)i * stride) * U) + 4 * lane;
            int r[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { const uint32_t row0 = hash_row(env + e, span); r[e] = (row0 % a.P) * a.M + row0 / a.P; }
            gather4(dst + lane * grp_pitch, &map, 0, r[0], r[1], r[2], r[3], bar);
        }
    };
    for (int i = 0; i < S && i < n_mine; ++i) issue(i);
    long long tc[6] = {0, 0, 0, 0, 0, 0};
    const int rows = U * W;
    // row r = lane + 32u of every unit: byte offset of its position-feature slot and its env's feature increment
    constexpr int kRounds = (U * 128 + 31) / 32;   // W <= 128
    int off_u[kRounds]; float pfe_u[kRounds];
#pragma unroll
    for (int u = 0; u < kRounds; ++u) {
        const int r = lane + 32 * u;
        const int e = r / W, g = e >> 2;
        off_u[u] = r < rows ? (int)(g * grp_pitch + (5 * (r - g * 4 * W) + 4) * 4) : -1;
        pfe_u[u] = (float)e * 0.001f;
    }
    for (int i = 0; i < n_mine; ++i) {
        const int s = i % S;
        const int64_t env0 = (first + (int64_t)i * stride) * U;
        long long t0 = clock64();
        if (!mbar_wait_bounded(bars + 8 * s, (i / S) & 1)) { if (lane == 0) atomicAdd(a.err, 1); return; }
        long long t1 = clock64(); tc[0] += t1 - t0;
        unsigned char *o = ring + (size_t)s * pitch;
        const float pf0 = (float)(uint32_t)env0 * 0.001f;
#pragma unroll
        for (int u = 0; u < kRounds; ++u)
            if (off_u[u] >= 0) *reinterpret_cast<float *>(o + off_u[u]) = pf0 + pfe_u[u];
        long long t2 = clock64(); tc[1] += t2 - t1;
        fence_async();
        __syncwarp();
        long long t3 = clock64(); tc[2] += t3 - t2;
        if (lane < G && !(a.flags & 1)) {
            bulk_store(a.obs + ((size_t)env0 + 4 * lane) * W * 5, smem_u32(o) + lane * grp_pitch, grp_b);
            bulk_commit();
        }
        long long t4 = clock64(); tc[3] += t4 - t3;
        const int j = i - (a.lag - 1);
        if (j >= 0 && j + S < n_mine) {
            if (lane < G) { if (a.lag == 1) bulk_wait_read<0>(); else if (a.lag == 2) bulk_wait_read<1>(); else bulk_wait_read<2>(); }
            __syncwarp();
            long long t5 = clock64(); tc[4] += t5 - t4;
            issue(j + S);
            tc[5] += clock64() - t5;
        }
    }
    if (lane < G) bulk_wait_read<0>();
    if (blockIdx.x == 0 && threadIdx.x == 0) { for (int k = 0; k < 6; ++k) a.clk[k] = (unsigned long long)tc[k]; a.clk[6] = (unsigned long long)n_mine; }
}

__global__ void check_hash_kernel(const Args a, unsigned long long *bad) {
    const int64_t n = a.N * a.W * 5;
    const uint32_t span = (uint32_t)(a.T - a.W);
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < n; f += (int64_t)gridDim.x * blockDim.x) {
        const int64_t env = f / (a.W * 5);
        const int rem = (int)(f - env * a.W * 5), j = rem / 5, c = rem % 5;
        const int64_t unit_envs = a.lag < 0 ? 1 : (int64_t)a.S;   // a.S carries the unit size for the check
        const int64_t env0 = env / unit_envs * unit_envs;
        const float want = c == 4 ? (float)(uint32_t)env0 * 0.001f + (float)(int)(env - env0) * 0.001f
                                  : a.plain[((size_t)hash_row((uint32_t)env, span) + j) * 4 + c];
        if (a.obs[f] != want) atomicAdd(bad, 1ull);
    }
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


template <typename L>
void run_generic(const char *name, Args a, int sms, int flags, unsigned long long *bad_dev, L launch) {
    a.flags = flags;
    CHECK(cudaMemset(a.err, 0, sizeof(int)));
    CHECK(cudaMemset(a.obs, 0xFF, (size_t)a.N * a.W * 20));
    launch(a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s flags %d: %s\n", name, flags, cudaGetErrorString(e)); exit(2); }
    int err = 0;
    CHECK(cudaMemcpy(&err, a.err, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) { printf("%s flags %d: %d warps timed out\n", name, flags, err); return; }
    unsigned long long bad = 0;
    if (flags == 0) {
        CHECK(cudaMemset(bad_dev, 0, 8));
        check_kernel<<<sms * 8, 256>>>(a, bad_dev);
        CHECK(cudaMemcpy(&bad, bad_dev, 8, cudaMemcpyDeviceToHost));
    }
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch(a);
    CHECK(cudaEventRecord(e0));
    const int reps = 20;
    for (int i = 0; i < reps; ++i) launch(a);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    printf("%s  flags %d   %.4f ms   out %7.1f GB/s   mismatches %llu\n", name, flags, ms, (double)a.N * a.W * 20 / ms / 1e6, bad);
    fflush(stdout);
}

template <int G>
void run_a2(const CUtensorMap &map, Args a, int sms, int nw, int S, int lag, int flags, unsigned long long *bad_dev) {
    const uint32_t pitch = G * ((4 * 20u * a.W + 127) & ~127u);
    const size_t smem = (size_t)nw * (128 + (size_t)S * pitch);
    if (smem > 226 * 1024) return;
    a.S = S; a.lag = lag;
    CHECK(cudaFuncSetAttribute(probe_a2_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    char name[128];
    snprintf(name, sizeof name, "A2 unit %2d envs  warps %2d  S %d  lag %d  smem %6zu", 4 * G, nw, S, lag, smem);
    run_generic(name, a, sms, flags, bad_dev, [&](const Args &aa) { probe_a2_kernel<G><<<sms, nw * 32, smem>>>(map, aa); });
}

void run_b(const CUtensorMap &map, Args a, int sms, int C, int S, int flags, unsigned long long *bad_dev) {
    const size_t smem = 256 + 8 * 32 * 4 + (size_t)S * 8 * ((4 * 20u * a.W + 127) & ~127u);
    if (smem > 226 * 1024 || S > 8) return;
    a.S = S;
    CHECK(cudaFuncSetAttribute(probe_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    char name[128];
    snprintf(name, sizeof name, "B  tile 32 envs  consumers %d  S %d  smem %6zu", C, S, smem);
    run_generic(name, a, sms, flags, bad_dev, [&](const Args &aa) { probe_b_kernel<<<sms, (C + 2) * 32, smem>>>(map, aa, C); });
}

template <int G>
void run_a3(const CUtensorMap &map, Args a, int sms, int nw, int S, int lag, int flags, unsigned long long *bad_dev) {
    const uint32_t pitch = G * ((4 * 20u * a.W + 127) & ~127u);
    const size_t smem = (size_t)nw * (128 + (size_t)S * pitch);
    if (smem > 226 * 1024) return;
    a.S = S; a.lag = lag; a.flags = flags;
    CHECK(cudaFuncSetAttribute(probe_a3_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CHECK(cudaMemset(a.err, 0, sizeof(int)));
    CHECK(cudaMemset(a.obs, 0xFF, (size_t)a.N * a.W * 20));
    auto launch = [&] { probe_a3_kernel<G><<<sms, nw * 32, smem>>>(map, a); };
    launch();
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("A3 G %d nw %d: %s\n", G, nw, cudaGetErrorString(e)); exit(2); }
    int err = 0;
    CHECK(cudaMemcpy(&err, a.err, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) { printf("A3 G %d nw %d S %d: %d warps timed out\n", G, nw, S, err); return; }
    unsigned long long bad = 0;
    if (flags == 0) {
        CHECK(cudaMemset(bad_dev, 0, 8));
        Args ac = a; ac.S = 4 * G;
        check_hash_kernel<<<sms * 8, 256>>>(ac, bad_dev);
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
    unsigned long long h[7];
    CHECK(cudaMemcpy(h, a.clk, 56, cudaMemcpyDeviceToHost));
    const double n = (double)h[6];
    printf("A3 unit %2d envs  warps %2d  S %d  lag %d  flags %d  smem %6zu  %.4f ms  %7.1f GB/s  bad %llu | clk/unit: wait %5.0f pf %4.0f fence %4.0f store %4.0f wgroup %5.0f issue %5.0f\n",
           4 * G, nw, S, lag, flags, smem, ms, (double)a.N * a.W * 20 / ms / 1e6, bad, h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[5] / n);
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
    a.T = T; CHECK(cudaMalloc(&a.clk, 64));
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

    const bool v1 = argc > 4 && atoi(argv[4]) == 1;
    if (v1) {
        run<0>(map_inter, a, sms, 8, 4, 2, 0, d_bad);
        run<1>(map_inter, a, sms, 8, 4, 2, 0, d_bad);
        run<2>(map_plain, a, sms, 8, 4, 2, 0, d_bad);
        run<3>(map_plain, a, sms, 8, 4, 2, 0, d_bad);
        for (int nw : {8, 16}) { run<0>(map_inter, a, sms, nw, 2, 1, 0, d_bad); run<2>(map_plain, a, sms, nw, 2, 1, 0, d_bad); }
    }
    // v3: lanes < G issue their group's load and store; phase clocks
    for (int nw : {8, 12, 16, 20, 24}) for (int S : {2, 3}) for (int lag : {1, 2}) run_a3<1>(map_inter, a, sms, nw, S, lag, 0, d_bad);
    for (int nw : {6, 8, 10, 11}) for (int S : {2}) for (int lag : {1, 2}) run_a3<2>(map_inter, a, sms, nw, S, lag, 0, d_bad);
    for (int nw : {4, 5, 7}) for (int S : {2, 3}) for (int lag : {1, 2}) run_a3<2>(map_inter, a, sms, nw, S, lag, 0, d_bad);
    for (int nw : {3, 4, 5}) for (int S : {2}) for (int lag : {1, 2}) run_a3<4>(map_inter, a, sms, nw, S, lag, 0, d_bad);
    for (int nw : {2}) for (int S : {2}) for (int lag : {1, 2}) run_a3<8>(map_inter, a, sms, nw, S, lag, 0, d_bad);
    run_a3<1>(map_inter, a, sms, 16, 2, 1, 1, d_bad); run_a3<1>(map_inter, a, sms, 16, 2, 1, 2, d_bad);
    run_a3<2>(map_inter, a, sms, 10, 2, 1, 1, d_bad); run_a3<2>(map_inter, a, sms, 10, 2, 1, 2, d_bad);
    run_a3<4>(map_inter, a, sms, 5, 2, 1, 1, d_bad); run_a3<4>(map_inter, a, sms, 5, 2, 1, 2, d_bad);
    return 0;
}
