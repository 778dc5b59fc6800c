// fe_step.cu — the vectorised trading-env step of hmomin/FinEnvs as hand-written sm_100a CUDA.
//
// One launch per step does what finenvs/environments/time_series_env.py:277-536 does in ~660
// torch-eager ops: advance the time pointer, read the current OHLC bar, map the action to trades
// with cash / position / margin / commission bookkeeping, build the observation window, compute
// the reward, detect episode end, auto-reset (with counter-based redraws) and keep the
// evaluate-mode metrics.  The arithmetic follows the reference op for op, dtype for dtype
// (SURVEY.md App. A): f32 state, f64 temporaries rounded once, no FMA contraction (every
// product/sum below is an explicit __*_rn intrinsic, and the TU is built with -fmad=false).
//
// Data movement (the part that costs time: ~2.2 KB per env-step at W=60, <100 flops):
//   tile variant   — each env's W x 16 B log-return window is fetched with one 1-D bulk async
//                    copy (cp.async.bulk, SASS UBLKCP) into shared memory, completing on an
//                    mbarrier; the block interleaves the position feature (4 -> 5 values per row,
//                    conflict-free stride-5 STS) into an output tile that is contiguous in the
//                    (N, W, 5) observation tensor and leaves with one bulk async store.
//   direct variant — warp-per-env global->global copy for windows that do not fit in smem.
//
// The C ABI is declared in include/finenvs_b200.h.
#include "finenvs_b200.h"
#include "fe_common.cuh"

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace {

constexpr int kThreads = 128;          // threads per block (both variants)
constexpr int kMaxTileEnvs = kThreads; // one bookkeeping thread per env of the tile
constexpr int kSmemHeader = 16;        // mbarrier (8 B) + pad
constexpr int kSmemMax = 226 * 1024;

// ------------------------------------------------------------------------------------------
// exact (never contracted) arithmetic helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float d2f(double a) { return __double2float_rn(a); }
// torch.relu keeps NaN (clamp_min), unlike fmax
__device__ __forceinline__ float relu32(float x) { return x < 0.0f ? 0.0f : x; }
__device__ __forceinline__ double relu64(double x) { return x < 0.0 ? 0.0 : x; }

struct Consts {
    float ms, scale, cf, imrf, SBf;
    double c, imr, mmr1, SB;
    // optional second destination of rewards / dones (fe_step_host's zero-copy mode: mapped pinned host memory)
    void *rewards_mirror;
    int32_t *dones_mirror;
};

Consts make_consts(const FeParams &p) {
    Consts k;
    k.ms = (float)p.max_shares;
    k.scale = (float)((double)p.max_shares + 0.5); // :299  f32 tensor * python float
    k.cf = (float)p.commission;                    // :364
    k.imrf = (float)p.imr;                         // :377-378
    k.SBf = (float)p.starting_balance;             // :499
    k.c = p.commission;
    k.imr = p.imr;
    k.mmr1 = 1.0 + p.mmr;                          // :462
    k.SB = p.starting_balance;
    k.rewards_mirror = nullptr;
    k.dones_mirror = nullptr;
    return k;
}

__device__ __forceinline__ void draw_segment(const FeParams &p, const FeSeries &s, int64_t gid, uint64_t step,
                                             uint32_t kind, int32_t &seg, int32_t &off) {
    uint32_t r[4];
    philox4x32_10(p.seed, (uint64_t)gid, step, kind, r);
    seg = (int32_t)__umulhi(r[0], (uint32_t)p.num_segments);
    off = 0;
    if (p.random_offset) {
        // valid start pointers are 0 .. seg_len - W - 1 (one bar must remain to step onto)
        const int32_t span = __ldg(s.seg_len + seg) - p.window;
        off = span > 0 ? (int32_t)__umulhi(r[1], (uint32_t)span) : 0;
    }
}

// ------------------------------------------------------------------------------------------
// per-env bookkeeping: everything of step() except moving the window
// ------------------------------------------------------------------------------------------
struct EnvResult {
    int64_t row0;   // first row of the observation window in the flat series
    double posfeat; // (long - short) * close / starting_balance  (:428-431)
    int done;
    int newly_terminated;
    double fin_return; // episode return of an env that finished this step (stats)
    int fin_len;
};

// reset() (:423-435): no state change, current window + position feature
__device__ __forceinline__ EnvResult env_observe(const FeParams &p, const FeSeries &s, const FeState &st,
                                                 const Consts &k, int64_t i) {
    EnvResult r;
    r.row0 = __ldg(s.seg_start + st.seg[i]) + st.ptr[i];
    const double C = __ldg(s.prices + (r.row0 + p.window - 1) * 4 + 3);
    const float net = fsub(st.long_sh[i], st.short_sh[i]);
    r.posfeat = __ddiv_rn(dmul((double)net, C), k.SB);
    r.done = 0; r.newly_terminated = 0; r.fin_return = 0.0; r.fin_len = 0;
    return r;
}

template <typename OutT>
__device__ __forceinline__ EnvResult env_step(const FeParams &p, const FeSeries &s, const FeState &st,
                                              const Consts &k, int64_t i, const float *__restrict__ actions,
                                              OutT *__restrict__ rewards, int32_t *__restrict__ dones,
                                              bool track, uint64_t step) {
    EnvResult res;
    const int W = p.window;
    // :298-302 action -> integer share delta (round half to even, then clamp)
    float d = rintf(fmul(__ldg(actions + i), k.scale));
    d = d < -k.ms ? -k.ms : (d > k.ms ? k.ms : d);
    // :281-282 advance time
    int32_t seg = st.seg[i];
    int32_t ptr = st.ptr[i] + 1;
    const int32_t len = __ldg(s.seg_len + seg);
    res.row0 = __ldg(s.seg_start + seg) + ptr;
    // :323-342 current bar = last row of the window: O,H,L,C as two 16-byte loads
    const double2 *px = reinterpret_cast<const double2 *>(s.prices + (res.row0 + W - 1) * 4);
    const double2 oh = __ldg(px), lc = __ldg(px + 1);
    const double O = oh.x, H = oh.y, L = lc.x, C = lc.y;
    float cash = st.cash[i];
    float lng = st.long_sh[i];
    float sht = st.short_sh[i];
    double margin = st.margin[i];
    float comm = 0.0f; // :305
    // :344-351
    float pos = d < 0.0f ? 0.0f : d;
    float neg = d > 0.0f ? 0.0f : d;
    const double Omc = dsub(O, k.c), Opc = dadd(O, k.c);
    { // :353-361 sell longs (+ :363-365)
        const float nl = relu32(fadd(lng, neg));
        const float sold = fsub(lng, nl);
        neg = fadd(neg, sold);
        comm = fadd(comm, fmul(sold, k.cf));
        cash = d2f(dadd((double)cash, dmul((double)sold, Omc)));
        lng = nl;
    }
    { // :367-383 cover shorts, re-mark margin to imr * short * open
        const float ns = relu32(fsub(sht, pos));
        const float bought = fsub(sht, ns);
        pos = fsub(pos, bought);
        comm = fadd(comm, fmul(bought, k.cf));
        cash = d2f(dsub((double)cash, dmul((double)bought, Opc)));
        sht = ns;
        const double nm = dmul((double)fmul(k.imrf, sht), O);
        cash = d2f(dsub((double)cash, dsub(nm, margin)));
        margin = nm;
    }
    // :385-392 all-or-nothing long entry
    if (dsub((double)cash, dmul((double)pos, Opc)) < 0.0) pos = 0.0f;
    // :394-399
    comm = fadd(comm, fmul(pos, k.cf));
    cash = d2f(dsub((double)cash, dmul((double)pos, Opc)));
    lng = fadd(lng, pos);
    { // :401-410 all-or-nothing short entry, :412-421 open short
        float q = -neg;
        float sc = fmul(q, k.cf);
        double req = dmul(k.imr, dmul((double)q, O));
        if (dsub(dsub((double)cash, req), (double)sc) < 0.0) {
            q = -0.0f; // neg = 0 -> -neg
            sc = fmul(q, k.cf);
            req = dmul(k.imr, dmul((double)q, O));
        }
        comm = fadd(comm, fmul(q, k.cf));
        cash = d2f(dsub((double)cash, dadd(req, (double)sc)));
        margin = dadd(margin, req);
        sht = fadd(sht, q);
    }
    // :321 -> reset(): the observation is built NOW, before rewards / dones / auto-reset
    res.posfeat = __ddiv_rn(dmul((double)fsub(lng, sht), C), k.SB);
    // :447-457 rewards
    int done = cash < 0.0f; // :448
    double rew;
    {
        // :459-468 maintenance margin at High
        const double mc1 = relu64(dsub(dmul(dmul((double)sht, H), k.mmr1), margin));
        cash = d2f(dsub((double)cash, mc1));
        margin = dadd(margin, mc1);
        done |= cash < 0.0f;
        // :470-475 margin release at Low
        const double rel = relu64(dsub(margin, dmul(dmul((double)sht, L), k.imr)));
        margin = dsub(margin, rel);
        cash = d2f(dadd((double)cash, rel));
        // :451 maintenance margin at Close
        const double mc2 = relu64(dsub(dmul(dmul((double)sht, C), k.mmr1), margin));
        cash = d2f(dsub((double)cash, mc2));
        margin = dadd(margin, mc2);
        done |= cash < 0.0f;
        rew = dadd(-mc1, -mc2);
        if (done) { lng = 0.0f; sht = 0.0f; } // :452-453
        rew = dadd(rew, dmul((double)fsub(lng, sht), dsub(C, O))); // :454-455
        rew = dsub(rew, (double)comm);                              // :456
    }
    // :477-496 time limit / NaN padding == pointer ran into the end of the (effective) segment
    done |= (ptr + W >= len);
    // :288-289 closing commission on whatever is still held
    rew = dsub(rew, (double)fmul(fmul(done ? 1.0f : 0.0f, fadd(sht, lng)), k.cf));
    // :498-521 auto-reset
    if (done) {
        cash = k.SBf; margin = 0.0; lng = 0.0f; sht = 0.0f; ptr = 0;
        const int64_t gid = p.env_id_base + i;
        if (p.reset_mode == FE_RESET_ALL || (p.reset_mode == FE_RESET_LAST && gid == p.total_envs - 1)) {
            draw_segment(p, s, gid, step, 0u, seg, ptr);
            st.seg[i] = seg;
        }
    }
    res.done = done;
    res.newly_terminated = 0; res.fin_return = 0.0; res.fin_len = 0;
    if (p.evaluate) { // :523-536
        const int was = st.terminated[i];
        if (was) rew = 0.0;                                             // :527-528
        if (done && !was) { st.terminated[i] = 1; res.newly_terminated = 1; } // :529
        st.ep_return[i] = d2f(dadd((double)st.ep_return[i], rew));      // :530  f32 += f64
    } else if (track) { // extension: running episode return / length for the NCCL-reduced statistics
        const float er = d2f(dadd((double)st.ep_return[i], rew));
        const int32_t el = st.ep_len[i] + 1;
        if (done) { res.fin_return = (double)er; res.fin_len = el; }
        st.ep_return[i] = done ? 0.0f : er;
        st.ep_len[i] = done ? 0 : el;
    }
    st.ptr[i] = ptr; st.cash[i] = cash; st.long_sh[i] = lng; st.short_sh[i] = sht; st.margin[i] = margin;
    rewards[i] = (OutT)rew;
    dones[i] = done; // :296 dones.int()
    if (k.dones_mirror) { reinterpret_cast<OutT *>(k.rewards_mirror)[i] = (OutT)rew; k.dones_mirror[i] = done; }
    return res;
}

// one atomic per warp for the episode statistics
__device__ __forceinline__ void accumulate_stats(FeStats *stats, const EnvResult &r, bool active) {
    if (stats == nullptr) return;
    const unsigned full = 0xFFFFFFFFu;
    const int done = active ? r.done : 0;
    const int nterm = active ? r.newly_terminated : 0;
    const unsigned nd = __reduce_add_sync(full, (unsigned)done);
    const unsigned nt = __reduce_add_sync(full, (unsigned)nterm);
    if (nd == 0 && nt == 0) return;
    unsigned sl = __reduce_add_sync(full, (unsigned)(done ? r.fin_len : 0));
    double sr = done ? r.fin_return : 0.0;
    double sq = sr * sr;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(full, sr, o);
        sq += __shfl_xor_sync(full, sq, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (nd) atomicAdd(&stats->n_done, (unsigned long long)nd);
        if (nt) atomicAdd(&stats->n_terminated, (unsigned long long)nt);
        if (sl) atomicAdd(&stats->sum_len, (unsigned long long)sl);
        if (nd) { atomicAdd(&stats->sum_return, sr); atomicAdd(&stats->sum_return_sq, sq); }
    }
}

// ------------------------------------------------------------------------------------------
// tile variant
// ------------------------------------------------------------------------------------------
template <typename OutT> struct Row4;
template <> struct Row4<float> { float4 v; };
template <> struct Row4<double> { double2 a, b; };

// smem layout: [mbarrier 16 B][posfeat E x OutT, padded to 16 B][in tile E*W*4 OutT][out tile E*W*5 OutT]
template <typename OutT> __host__ __device__ constexpr size_t tile_pf_bytes(int E) {
    return ((size_t)E * sizeof(OutT) + 15) & ~(size_t)15;
}
template <typename OutT> __host__ __device__ inline size_t tile_smem_bytes(int E, int W) {
    return kSmemHeader + tile_pf_bytes<OutT>(E) + (size_t)E * W * 9 * sizeof(OutT);
}

template <typename OutT, bool kObserve>
__global__ void __launch_bounds__(kThreads)
fe_tile_kernel(const FeParams p, const FeSeries s, const FeState st, const Consts k, const float *__restrict__ actions,
               OutT *__restrict__ obs, OutT *__restrict__ rewards, int32_t *__restrict__ dones, FeStats *stats,
               const uint64_t step_arg, const uint64_t *__restrict__ step_dev, const int E) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint64_t step = step_dev ? *step_dev : step_arg;
    const int W = p.window;
    const int tid = threadIdx.x;
    const int64_t env0 = (int64_t)blockIdx.x * E;
    const int nvalid = (int)min((int64_t)E, p.num_envs - env0);
    OutT *pf = reinterpret_cast<OutT *>(smem + kSmemHeader);
    unsigned char *in_tile = smem + kSmemHeader + tile_pf_bytes<OutT>(E);
    OutT *out_tile = reinterpret_cast<OutT *>(in_tile + (size_t)E * W * 4 * sizeof(OutT));
    const uint32_t bar = smem_u32(smem);
    const uint32_t row_bytes = 4 * sizeof(OutT);
    const uint32_t win_bytes = (uint32_t)W * row_bytes;

    if (tid == 0) {
        mbar_init(bar, (uint32_t)nvalid);
        mbar_fence_init();
    }
    __syncthreads();

    // ---- one thread per env: fetch its window asynchronously, do the bookkeeping meanwhile ----
    const bool active = tid < nvalid;
    EnvResult r;
    r.done = 0; r.newly_terminated = 0; r.fin_return = 0.0; r.fin_len = 0;
    if (active) {
        const int64_t i = env0 + tid;
        // window start is known before any arithmetic: launch the copy first
        const int64_t row0 = __ldg(s.seg_start + st.seg[i]) + st.ptr[i] + (kObserve ? 0 : 1);
        mbar_arrive_expect_tx(bar, win_bytes);
        bulk_load(smem_u32(in_tile + (size_t)tid * win_bytes),
                  reinterpret_cast<const unsigned char *>(s.logret) + (size_t)row0 * row_bytes, win_bytes, bar);
        if (kObserve) r = env_observe(p, s, st, k, i);
        else r = env_step<OutT>(p, s, st, k, i, actions, rewards, dones, stats != nullptr, step);
        pf[tid] = (OutT)r.posfeat;
    }
    if (!kObserve && tid < ((nvalid + 31) & ~31)) accumulate_stats(stats, r, active);
    __syncthreads(); // pf visible
    mbar_wait(bar, 0); // all windows landed

    // ---- interleave: row (4 values) + position feature -> 5 values, rows are contiguous in both tiles
    const int nrows = nvalid * W;
    const float invW = 1.0f / (float)W;
    const Row4<OutT> *in_rows = reinterpret_cast<const Row4<OutT> *>(in_tile);
    for (int row = tid; row < nrows; row += blockDim.x) {
        const int e = __float2int_rz(((float)row + 0.5f) * invW);
        const Row4<OutT> v = in_rows[row];
        OutT *o = out_tile + (size_t)row * 5;
        if constexpr (sizeof(OutT) == 4) {
            o[0] = v.v.x; o[1] = v.v.y; o[2] = v.v.z; o[3] = v.v.w;
        } else {
            o[0] = v.a.x; o[1] = v.a.y; o[2] = v.b.x; o[3] = v.b.y;
        }
        o[4] = pf[e];
    }
    // ---- out tile -> obs[env0 : env0+nvalid] (contiguous): one bulk async store
    const size_t out_bytes = (size_t)nrows * 5 * sizeof(OutT);
    OutT *dst = obs + (size_t)env0 * W * 5;
    if ((out_bytes & 15) == 0) {
        fence_proxy_async_smem(); // generic-proxy smem writes -> visible to the async proxy
        __syncthreads();
        if (tid == 0) {
            bulk_store(dst, smem_u32(out_tile), (uint32_t)out_bytes);
            bulk_commit();
            bulk_wait_read_all(); // smem must outlive the read side of the store
        }
    } else { // ragged tail block whose byte count is not a multiple of 16
        __syncthreads();
        for (int f = tid; f < nrows * 5; f += blockDim.x) dst[f] = out_tile[f];
    }
}

// ------------------------------------------------------------------------------------------
// pipe variant: persistent, warp-specialised (the fast path; windows up to 512 rows)
//
// One block per SM loops over tiles of TE consecutive envs.  Roles:
//   bookkeeper warps (kPipeBook): one lane per env at full lane utilisation; run the per-env arithmetic and
//       publish {row0, position feature} of a tile into a ring of Q descriptors (mbarriers desc_full /
//       desc_free).  They run up to Q tiles ahead, so their load -> lookup -> load -> arithmetic latency
//       chain is off the critical path.
//   mover warps (kPipeMove): gather the tile's window rows with coalesced 16-byte loads straight into
//       registers (RPT rows per thread, the loads of tile t+1 are issued before tile t is written, so
//       their L2/HBM latency hides behind a whole tile of work), interleave the position feature while
//       storing into the out ring (conflict-free stride-5 STS), and one thread issues ONE bulk async store
//       (UBLKCP.G.S) per tile; up to S_OUT stores stay in flight.
// History (profiles/r01_pipe_*.txt): a first version fetched every env's window with its own bulk async
// copy into an in-ring; ncu showed the block pinned on the copy-issue loop — one UBLKCP per ~85 cycles
// per SM whatever the stage counts (7085 copies per SM per step => 0.32 ms floor), the same wall the
// tile variant hits.  Register gathers have no such per-copy cost and halve the shared-memory traffic.
// ------------------------------------------------------------------------------------------
#ifndef FE_PIPE_BOOK
#define FE_PIPE_BOOK 6
#endif
#ifndef FE_PIPE_MOVE
#define FE_PIPE_MOVE 8
#endif
#ifndef FE_PIPE_SOUT
#define FE_PIPE_SOUT 2
#endif
#ifndef FE_PIPE_SIN
#define FE_PIPE_SIN 4   /* in-ring stages of the "stream" flavour (series larger than L2) */
#endif
#ifndef FE_PIPE_RPT
#define FE_PIPE_RPT 8   /* f32 rows per mover thread per tile (f64: half) */
#endif
constexpr int kPipeBook = FE_PIPE_BOOK;
constexpr int kPipeMove = FE_PIPE_MOVE;
constexpr int kPipeThreads = (kPipeBook + kPipeMove) * 32;
constexpr int kMovers = kPipeMove * 32;
constexpr int kPipeQ = 8;                // descriptor ring depth
constexpr int kPipeSOut = FE_PIPE_SOUT;  // out-tile stages
// Two flavours, chosen per launch from the size of the log-return table (pick_pipe_stages):
//   kSIn == 0  "cached": the table is L2-resident, one tile of register prefetch covers the L2 latency and the
//              rows never touch shared memory on the way in (measured c2: 0.28 ms vs 0.35 ms for kSIn == 3);
//   kSIn  > 0  "stream": the table lives in HBM; rows arrive through an in-ring of kSIn stages filled with
//              16-byte cp.async (LDGSTS), ~kSIn x 30 KB in flight per SM (measured c4: 0.376 ms vs 0.43 ms).
constexpr int kPipeSInStream = FE_PIPE_SIN;
template <typename OutT> struct PipeRows { static constexpr int value = sizeof(OutT) == 4 ? FE_PIPE_RPT : (FE_PIPE_RPT + 1) / 2; }; // rows / thread / tile

// smem: [mbarriers desc_full[Q], desc_free[Q], in_full[S_IN]] (256 B) [descriptors Q x TE x (8 + 8) B]
//       [in ring S_IN x TE*W*4 OutT] [out ring S_OUT x TE*W*5 OutT]
template <typename OutT> __host__ __device__ inline size_t pipe_smem_bytes(int TE, int W, int sin) {
    return 256 + (size_t)kPipeQ * TE * 16 + (size_t)TE * W * sizeof(OutT) * (4 * sin + 5 * kPipeSOut);
}
// 16-byte async copy global -> shared (LDGSTS), and "arrive on the mbarrier once my copies have landed"
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ Row4<float> ldg_row(const Row4<float> *p) {
    Row4<float> r;
    r.v = __ldg(reinterpret_cast<const float4 *>(p));
    return r;
}
__device__ __forceinline__ Row4<double> ldg_row(const Row4<double> *p) {
    Row4<double> r;
    r.a = __ldg(reinterpret_cast<const double2 *>(p));
    r.b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    return r;
}

template <typename OutT, bool kObserve, int kPipeSIn>
__global__ void __launch_bounds__(kPipeThreads, 1)
fe_pipe_kernel(const FeParams p, const FeSeries s, const FeState st, const Consts k, const float *__restrict__ actions,
               OutT *__restrict__ obs, OutT *__restrict__ rewards, int32_t *__restrict__ dones, FeStats *stats,
               const uint64_t step_arg, const uint64_t *__restrict__ step_dev, const int TE) {
    constexpr int RPT = PipeRows<OutT>::value;
    extern __shared__ __align__(128) unsigned char smem[];
    const uint64_t step = step_dev ? *step_dev : step_arg;
    const int W = p.window;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t ntiles_all = (p.num_envs + TE - 1) / TE;
    const int ntiles = (int)((ntiles_all - blockIdx.x + gridDim.x - 1) / gridDim.x); // tiles blockIdx.x, +gridDim.x, ...
    const uint32_t bars = smem_u32(smem);
    auto desc_full = [&](int q) { return bars + 8u * q; };
    auto desc_free = [&](int q) { return bars + 8u * (kPipeQ + q); };
    auto in_full = [&](int si) { return bars + 8u * (2 * kPipeQ + si); };
    int64_t *d_row0 = reinterpret_cast<int64_t *>(smem + 256);                       // [Q][TE]
    double *d_pf = reinterpret_cast<double *>(smem + 256 + (size_t)kPipeQ * TE * 8); // [Q][TE], OutT in the low bytes
    const size_t in_stage = (size_t)TE * W * 4 * sizeof(OutT), out_stage = (size_t)TE * W * 5 * sizeof(OutT);
    unsigned char *in_ring = smem + 256 + (size_t)kPipeQ * TE * 16;
    unsigned char *out_ring = in_ring + kPipeSIn * in_stage;

    if (tid == 0) {
        for (int q = 0; q < kPipeQ; ++q) { mbar_init(desc_full(q), 1); mbar_init(desc_free(q), 1); }
        for (int si = 0; si < kPipeSIn; ++si) mbar_init(in_full(si), kMovers);
        mbar_fence_init();
    }
    __syncthreads();

    if (warp < kPipeBook) {
        // ------------------------------------------------------------------ bookkeepers
        for (int t = warp; t < ntiles; t += kPipeBook) {
            const int q = t % kPipeQ;
            mbar_wait(desc_free(q), ((t / kPipeQ) & 1) ^ 1); // first lap passes immediately
            const int64_t env0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TE;
            const int nvalid = (int)min((int64_t)TE, p.num_envs - env0);
            EnvResult r;
            r.done = 0; r.newly_terminated = 0; r.fin_return = 0.0; r.fin_len = 0; r.row0 = 0; r.posfeat = 0.0;
            const bool active = lane < nvalid;
            if (active) {
                const int64_t i = env0 + lane;
                if (kObserve) r = env_observe(p, s, st, k, i);
                else r = env_step<OutT>(p, s, st, k, i, actions, rewards, dones, stats != nullptr, step);
                d_row0[q * TE + lane] = r.row0;
                reinterpret_cast<OutT *>(d_pf + q * TE)[lane] = (OutT)r.posfeat;
            }
            if (!kObserve) accumulate_stats(stats, r, active);
            __syncwarp();
            if (lane == 0) mbar_arrive(desc_full(q)); // release: descriptor visible to the movers
        }
    } else {
        // ------------------------------------------------------------------ movers
        const int mtid = tid - kPipeBook * 32;
        int e_u[RPT], j_u[RPT]; // tile row mtid + u*kMovers = window row j_u of env e_u: the same in every tile
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
            e_u[u] = (mtid + u * kMovers) / W;
            j_u[u] = (mtid + u * kMovers) - e_u[u] * W;
        }
        const Row4<OutT> *series_rows = reinterpret_cast<const Row4<OutT> *>(s.logret);
        auto tile_rows = [&](int t) {
            const int64_t env0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TE;
            return (int)min((int64_t)TE, p.num_envs - env0) * W;
        };
        // fetch tile t's rows: row r of the tile = env r / W, window row r % W.  Either straight into registers
        // (kPipeSIn == 0) or with 16-byte async copies into in-ring stage t % S_IN, arriving on in_full when landed.
        auto gather = [&](int t, Row4<OutT>(&buf)[RPT]) {
            const int q = t % kPipeQ;
            mbar_wait(desc_full(q), (t / kPipeQ) & 1);
            const int nrows = tile_rows(t);
            const uint32_t stage = kPipeSIn > 0 ? smem_u32(in_ring + (t % (kPipeSIn > 0 ? kPipeSIn : 1)) * in_stage) : 0u;
#pragma unroll
            for (int u = 0; u < RPT; ++u) {
                const int r = mtid + u * kMovers;
                if (r < nrows) {
                    const Row4<OutT> *src = series_rows + d_row0[q * TE + e_u[u]] + j_u[u];
                    if constexpr (kPipeSIn == 0) {
                        buf[u] = ldg_row(src);
                    } else {
                        cp_async16(stage + (uint32_t)r * sizeof(Row4<OutT>), src);
                        if constexpr (sizeof(OutT) == 8)
                            cp_async16(stage + (uint32_t)r * sizeof(Row4<OutT>) + 16, reinterpret_cast<const char *>(src) + 16);
                    }
                }
            }
            if constexpr (kPipeSIn > 0) cp_async_arrive_noinc(in_full(t % (kPipeSIn > 0 ? kPipeSIn : 1)));
        };
        Row4<OutT> cur[RPT], nxt[RPT];
        constexpr int kAhead = kPipeSIn > 0 ? kPipeSIn : 1;
        if constexpr (kPipeSIn == 0) {
            if (ntiles > 0) gather(0, cur);
        } else {
            for (int t = 0; t < kAhead && t < ntiles; ++t) gather(t, nxt);
        }
        for (int t = 0; t < ntiles; ++t) {
            const int q = t % kPipeQ, so = t % kPipeSOut;
            const int64_t env0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TE;
            const int nrows = tile_rows(t);
            if constexpr (kPipeSIn == 0) {
                if (t + 1 < ntiles) gather(t + 1, nxt); // in flight while this tile is written
            } else {
                mbar_wait(in_full(t % kAhead), (t / kAhead) & 1);
                const Row4<OutT> *in_rows = reinterpret_cast<const Row4<OutT> *>(in_ring + (t % kAhead) * in_stage);
#pragma unroll
                for (int u = 0; u < RPT; ++u) {
                    const int r = mtid + u * kMovers;
                    if (r < nrows) cur[u] = in_rows[r];
                }
            }
            if (t >= kPipeSOut) { // the store that last used out[so] must have finished reading shared memory
                if (mtid == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPipeSOut - 1) : "memory");
                named_bar_sync(1, kMovers);
            }
            OutT *out_tile = reinterpret_cast<OutT *>(out_ring + so * out_stage);
            const OutT *pf = reinterpret_cast<const OutT *>(d_pf + q * TE);
#pragma unroll
            for (int u = 0; u < RPT; ++u) {
                const int r = mtid + u * kMovers;
                if (r < nrows) {
                    OutT *o = out_tile + (size_t)r * 5;
                    if constexpr (sizeof(OutT) == 4) {
                        o[0] = cur[u].v.x; o[1] = cur[u].v.y; o[2] = cur[u].v.z; o[3] = cur[u].v.w;
                    } else {
                        o[0] = cur[u].a.x; o[1] = cur[u].a.y; o[2] = cur[u].b.x; o[3] = cur[u].b.y;
                    }
                    o[4] = pf[e_u[u]];
                }
            }
            const size_t out_bytes = (size_t)nrows * 5 * sizeof(OutT);
            OutT *dst = obs + (size_t)env0 * W * 5;
            if ((out_bytes & 15) == 0) {
                fence_proxy_async_smem();
                named_bar_sync(1, kMovers); // out tile complete, descriptor consumed
                if (mtid == 0) {
                    bulk_store(dst, smem_u32(out_tile), (uint32_t)out_bytes);
                    bulk_commit();
                }
            } else { // ragged last tile
                named_bar_sync(1, kMovers);
                for (int f = mtid; f < nrows * 5; f += kMovers) dst[f] = out_tile[f];
                named_bar_sync(1, kMovers);
            }
            if constexpr (kPipeSIn == 0) {
                if (mtid == 0) mbar_arrive(desc_free(q));
#pragma unroll
                for (int u = 0; u < RPT; ++u) cur[u] = nxt[u];
            } else {
                // the barrier above also means every mover finished reading in-stage t % S_IN: refill it, and only
                // then release the descriptor (the refill of tile t+S_IN reads ITS descriptor, not this one)
                if (mtid == 0) mbar_arrive(desc_free(q));
                if (t + kAhead < ntiles) gather(t + kAhead, nxt);
            }
        }
        if (mtid == 0) bulk_wait_read_all();
    }
}

// ------------------------------------------------------------------------------------------
// scatter variant: the pipe variant without register staging.  ncu on the pipe kernel's "cached" flavour
// (profiles/r01_v3_pipe_ncu_full_c2.txt) shows nothing saturated — DRAM 58 %, L1->XBAR 63 %, LSU 40 % — while a plain
// fill of the same 1.26 GB runs at 7.4 TB/s: the movers' read side is bounded by what their registers can keep in
// flight (one tile, ~30 KB per SM).  Here the window elements travel global -> shared memory with element-sized
// asynchronous copies (cp.async, SASS LDGSTS) that land DIRECTLY at their interleaved position in the output tile
// (element i of an env's window goes to i + i/4: the 4 -> 5 interleave is the scatter's address pattern); the movers
// only add the position feature column.  Nothing passes through registers, so the bytes in flight are bounded by the
// shared-memory ring (kScDepth tiles ahead), not by the register file.  Roles per block (one block per SM):
//   bookkeeper warps: as in the pipe variant (descriptor ring of {row0, position feature});
//   mover warps: issue the copies of tile t, then retire tile t - depth (cp.async.wait_group, proxy fence, arrive);
//   one store warp: waits for a tile to be complete, issues its bulk async store (UBLKCP.G.S), frees the stage of the
//                   store before it once that one has been read out.
// ------------------------------------------------------------------------------------------
#ifndef FE_SC_BOOK
#define FE_SC_BOOK 6
#endif
#ifndef FE_SC_MOVE
#define FE_SC_MOVE 8
#endif
constexpr int kScBook = FE_SC_BOOK;
constexpr int kScMove = FE_SC_MOVE;
constexpr int kScThreads = (kScBook + kScMove + 1) * 32;
constexpr int kScMovers = kScMove * 32;
constexpr int kScQ = 8;        // descriptor ring depth
constexpr int kScMaxStages = 8;

template <typename OutT> __host__ __device__ inline size_t scatter_smem_bytes(int TE, int W, int stages) {
    return 256 + (size_t)kScQ * TE * 16 + (size_t)stages * (((size_t)TE * W * 5 * sizeof(OutT) + 127) & ~(size_t)127);
}
template <int kBytes> __device__ __forceinline__ void cp_async_elem(uint32_t dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst_smem), "l"(src), "n"(kBytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_dyn(int n) { // wait until at most n of this thread's groups are pending
    switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    }
}

template <typename OutT, bool kObserve>
__global__ void __launch_bounds__(kScThreads, 1)
fe_scatter_kernel(const FeParams p, const FeSeries s, const FeState st, const Consts k, const float *__restrict__ actions,
                  OutT *__restrict__ obs, OutT *__restrict__ rewards, int32_t *__restrict__ dones, FeStats *stats,
                  const uint64_t step_arg, const uint64_t *__restrict__ step_dev, const int TE, const int S, const int D) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint64_t step = step_dev ? *step_dev : step_arg;
    const int W = p.window;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t ntiles_all = (p.num_envs + TE - 1) / TE;
    const int ntiles = (int)((ntiles_all - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const uint32_t bars = smem_u32(smem);
    auto desc_full = [&](int q) { return bars + 8u * q; };
    auto desc_free = [&](int q) { return bars + 8u * (kScQ + q); };
    auto tile_ready = [&](int si) { return bars + 8u * (2 * kScQ + si); };
    auto stage_free = [&](int si) { return bars + 8u * (2 * kScQ + kScMaxStages + si); };
    int64_t *d_row0 = reinterpret_cast<int64_t *>(smem + 256);                      // [Q][TE]
    double *d_pf = reinterpret_cast<double *>(smem + 256 + (size_t)kScQ * TE * 8);  // [Q][TE], OutT in the low bytes
    const size_t stage_bytes = ((size_t)TE * W * 5 * sizeof(OutT) + 127) & ~(size_t)127;
    unsigned char *ring = smem + 256 + (size_t)kScQ * TE * 16;
    auto tile_env0 = [&](int t) { return ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TE; };

    if (tid == 0) {
        for (int q = 0; q < kScQ; ++q) { mbar_init(desc_full(q), 1); mbar_init(desc_free(q), kScMove); }
        for (int si = 0; si < kScMaxStages; ++si) { mbar_init(tile_ready(si), kScMove); mbar_init(stage_free(si), 1); }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp < kScBook) {
        // ------------------------------------------------------------------ bookkeepers
        for (int t = warp; t < ntiles; t += kScBook) {
            const int q = t % kScQ;
            mbar_wait(desc_free(q), ((t / kScQ) & 1) ^ 1); // first lap passes immediately
            const int64_t env0 = tile_env0(t);
            const int nvalid = (int)min((int64_t)TE, p.num_envs - env0);
            EnvResult r;
            r.done = 0; r.newly_terminated = 0; r.fin_return = 0.0; r.fin_len = 0; r.row0 = 0; r.posfeat = 0.0;
            const bool active = lane < nvalid;
            if (active) {
                const int64_t i = env0 + lane;
                if (kObserve) r = env_observe(p, s, st, k, i);
                else r = env_step<OutT>(p, s, st, k, i, actions, rewards, dones, stats != nullptr, step);
                d_row0[q * TE + lane] = r.row0;
                reinterpret_cast<OutT *>(d_pf + q * TE)[lane] = (OutT)r.posfeat;
            }
            if (!kObserve) accumulate_stats(stats, r, active);
            __syncwarp();
            if (lane == 0) mbar_arrive(desc_full(q));
        }
    } else if (warp < kScBook + kScMove) {
        // ------------------------------------------------------------------ movers
        const int mt = tid - kScBook * 32;
        const int W4 = W * 4, W5 = W * 5;
        const int step_e4 = kScMovers / W4, step_i4 = kScMovers % W4; // element walk: idx += kScMovers
        const int step_eW = kScMovers / W, step_jW = kScMovers % W;   // row walk
        const OutT *lr = reinterpret_cast<const OutT *>(s.logret);
        auto retire = [&](int tt, int pending) { // tile tt's copies of this thread have landed -> visible to the bulk store
            cp_async_wait_dyn(pending);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(tile_ready(tt % S));
        };
        for (int t = 0; t < ntiles; ++t) {
            const int q = t % kScQ, si = t % S;
            mbar_wait(desc_full(q), (t / kScQ) & 1);
            if (t >= S) mbar_wait(stage_free(si), ((t / S) - 1) & 1);
            const int nvalid = (int)min((int64_t)TE, p.num_envs - tile_env0(t));
            unsigned char *stage = ring + (size_t)si * stage_bytes;
            const uint32_t stage_u32 = smem_u32(stage);
            const int64_t *row0s = d_row0 + q * TE;
            { // window elements: element i of env e -> out element e*W5 + i + i/4
                int e = mt / W4, i = mt - e * W4;
                while (e < nvalid) {
                    const OutT *src = lr + row0s[e] * 4 + i;
                    cp_async_elem<sizeof(OutT)>(stage_u32 + (uint32_t)(e * W5 + i + (i >> 2)) * (uint32_t)sizeof(OutT), src);
                    e += step_e4; i += step_i4;
                    if (i >= W4) { i -= W4; ++e; }
                }
            }
            { // position feature column
                const OutT *pf = reinterpret_cast<const OutT *>(d_pf + q * TE);
                OutT *out = reinterpret_cast<OutT *>(stage);
                int e = mt / W, j = mt - e * W;
                while (e < nvalid) {
                    out[(e * W + j) * 5 + 4] = pf[e];
                    e += step_eW; j += step_jW;
                    if (j >= W) { j -= W; ++e; }
                }
            }
            cp_async_commit();
            __syncwarp();
            if (lane == 0) mbar_arrive(desc_free(q)); // this warp has consumed the descriptor
            if (t >= D) retire(t - D, D);
        }
        for (int tt = ntiles > D ? ntiles - D : 0; tt < ntiles; ++tt) retire(tt, ntiles - 1 - tt);
    } else {
        // ------------------------------------------------------------------ store warp
        for (int t = 0; t < ntiles; ++t) {
            const int si = t % S;
            mbar_wait(tile_ready(si), (t / S) & 1);
            const int64_t env0 = tile_env0(t);
            const int nvalid = (int)min((int64_t)TE, p.num_envs - env0);
            const size_t out_bytes = (size_t)nvalid * W * 5 * sizeof(OutT);
            OutT *dst = obs + (size_t)env0 * W * 5;
            const unsigned char *stage = ring + (size_t)si * stage_bytes;
            if ((out_bytes & 15) == 0) {
                if (lane == 0) {
                    bulk_store(dst, smem_u32(stage), (uint32_t)out_bytes);
                    bulk_commit();
                    if (t >= 1) { // at most this store still reading: the one before it has left shared memory
                        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        mbar_arrive(stage_free((t - 1) % S));
                    }
                }
            } else { // ragged last tile: plain stores, synchronous
                const OutT *src = reinterpret_cast<const OutT *>(stage);
                for (int f = lane; f < nvalid * W * 5; f += 32) dst[f] = src[f];
                __syncwarp();
                if (lane == 0) {
                    if (t >= 1) {
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        mbar_arrive(stage_free((t - 1) % S));
                    }
                }
            }
            __syncwarp();
        }
        if (lane == 0) bulk_wait_read_all();
    }
}

// ------------------------------------------------------------------------------------------
// direct variant: any window; thread-per-env bookkeeping, then warp-per-env copy
// ------------------------------------------------------------------------------------------
template <typename OutT, bool kObserve>
__global__ void __launch_bounds__(kThreads)
fe_direct_kernel(const FeParams p, const FeSeries s, const FeState st, const Consts k, const float *__restrict__ actions,
                 OutT *__restrict__ obs, OutT *__restrict__ rewards, int32_t *__restrict__ dones, FeStats *stats,
                 const uint64_t step_arg, const uint64_t *__restrict__ step_dev) {
    __shared__ int64_t sh_row0[kThreads];
    const uint64_t step = step_dev ? *step_dev : step_arg;
    __shared__ OutT sh_pf[kThreads];
    const int W = p.window;
    const int tid = threadIdx.x;
    const int64_t env0 = (int64_t)blockIdx.x * kThreads;
    const int nvalid = (int)min((int64_t)kThreads, p.num_envs - env0);
    const bool active = tid < nvalid;
    EnvResult r;
    r.done = 0; r.newly_terminated = 0; r.fin_return = 0.0; r.fin_len = 0; r.row0 = 0; r.posfeat = 0.0;
    if (active) {
        const int64_t i = env0 + tid;
        if (kObserve) r = env_observe(p, s, st, k, i);
        else r = env_step<OutT>(p, s, st, k, i, actions, rewards, dones, stats != nullptr, step);
    }
    sh_row0[tid] = r.row0;
    sh_pf[tid] = (OutT)r.posfeat;
    if (!kObserve) accumulate_stats(stats, r, active);
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    const OutT *lr = reinterpret_cast<const OutT *>(s.logret);
    const int nvals = W * 5;
    for (int e = warp; e < nvalid; e += kThreads / 32) {
        const OutT *src = lr + sh_row0[e] * 4;
        OutT *dst = obs + (size_t)(env0 + e) * nvals;
        const OutT pfe = sh_pf[e];
        for (int f = lane; f < nvals; f += 32) {
            const int j = f / 5, c = f - j * 5;
            dst[f] = c == 4 ? pfe : __ldg(src + j * 4 + c);
        }
    }
}

// ------------------------------------------------------------------------------------------
// lazy variant: the step WITHOUT materialising the observation.  An observation of this env is fully described by
// (row0, position feature): obs[i, j, 0:4] = logret[row0[i] + j], obs[i, j, 4] = posfeat[i] (:437-445, :428-434).
// Consumers that read the window straight from the staged series (the fused ES policy kernel, fe_es.cu) take the
// 12-byte handle instead of the W x 20-byte tensor; fe_materialize turns a handle into the tensor.
// ------------------------------------------------------------------------------------------
template <typename OutT, bool kObserve>
__global__ void __launch_bounds__(kThreads)
fe_lazy_kernel(const FeParams p, const FeSeries s, const FeState st, const Consts k, const float *__restrict__ actions,
               int64_t *__restrict__ row0_out, OutT *__restrict__ pf_out, OutT *__restrict__ rewards,
               int32_t *__restrict__ dones, FeStats *stats, const uint64_t step_arg, const uint64_t *__restrict__ step_dev) {
    const uint64_t step = step_dev ? *step_dev : step_arg;
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool active = i < p.num_envs;
    EnvResult r;
    r.done = 0; r.newly_terminated = 0; r.fin_return = 0.0; r.fin_len = 0; r.row0 = 0; r.posfeat = 0.0;
    if (active) {
        if (kObserve) r = env_observe(p, s, st, k, i);
        else r = env_step<OutT>(p, s, st, k, i, actions, rewards, dones, stats != nullptr, step);
        row0_out[i] = r.row0;
        pf_out[i] = (OutT)r.posfeat;
    }
    if (!kObserve) accumulate_stats(stats, r, active);
}

// handle -> tensor: warp per env, same element order as the direct variant
template <typename OutT>
__global__ void __launch_bounds__(kThreads)
fe_materialize_kernel(const int64_t N, const int W, const OutT *__restrict__ logret, const int64_t *__restrict__ row0,
                      const OutT *__restrict__ pf, OutT *__restrict__ obs) {
    const int lane = threadIdx.x & 31;
    const int64_t e = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
    if (e >= N) return;
    const OutT *src = logret + row0[e] * 4;
    OutT *dst = obs + (size_t)e * W * 5;
    const OutT pfe = pf[e];
    for (int f = lane; f < W * 5; f += 32) {
        const int j = f / 5, c = f - j * 5;
        dst[f] = c == 4 ? pfe : __ldg(src + j * 4 + c);
    }
}

// ------------------------------------------------------------------------------------------
// split variant: bookkeeping and streaming as two launches (what made the portfolio path 25 % faster).
//   fe_book_kernel    one THREAD per env at full occupancy runs env_step() and leaves {row0, position feature} as a
//                     header inside the env's slice of the observation tensor (row0 in its first 8 bytes, the feature
//                     at element 4 = its final place in window row 0);
//   fe_stream_kernel  one WARP per env, no shared memory, no barriers: lane l produces output elements l, l+32, ... of
//                     the env's W*5 values — element f is column f%5 of window row f/5, i.e. input element f - f/5 or
//                     the position feature — so every store instruction writes 128 contiguous bytes and every load
//                     instruction reads ~104 contiguous bytes of the staged series; ceil(5W/32) independent loads per
//                     lane are in flight at once and 48 warps per SM hide the L2 / HBM latency.
// ------------------------------------------------------------------------------------------
template <typename OutT, bool kObserve>
__global__ void __launch_bounds__(kThreads)
fe_book_kernel(const FeParams p, const FeSeries s, const FeState st, const Consts k, const float *__restrict__ actions,
               OutT *__restrict__ obs, OutT *__restrict__ rewards, int32_t *__restrict__ dones, FeStats *stats,
               const uint64_t step_arg, const uint64_t *__restrict__ step_dev) {
    const uint64_t step = step_dev ? *step_dev : step_arg;
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool active = i < p.num_envs;
    EnvResult r;
    r.done = 0; r.newly_terminated = 0; r.fin_return = 0.0; r.fin_len = 0; r.row0 = 0; r.posfeat = 0.0;
    if (active) {
        if (kObserve) r = env_observe(p, s, st, k, i);
        else r = env_step<OutT>(p, s, st, k, i, actions, rewards, dones, stats != nullptr, step);
        OutT *slice = obs + (size_t)i * p.window * 5;
        if constexpr (sizeof(OutT) == 8) {
            reinterpret_cast<int64_t *>(slice)[0] = r.row0;
        } else { // the slice is only 4-byte aligned when W is odd: two words
            reinterpret_cast<uint32_t *>(slice)[0] = (uint32_t)((uint64_t)r.row0 & 0xFFFFFFFFu);
            reinterpret_cast<uint32_t *>(slice)[1] = (uint32_t)((uint64_t)r.row0 >> 32);
        }
        slice[4] = (OutT)r.posfeat;
    }
    if (!kObserve) accumulate_stats(stats, r, active);
}

constexpr int kStreamThreads = 256;
constexpr int kStreamSlots = 15;                   // output elements per lane and batch whose loads are in flight together
constexpr int kStreamSpan = 32 * kStreamSlots;     // 480 output elements = 96 window rows: a multiple of 5, so the
constexpr int kStreamSpanIn = kStreamSpan / 5 * 4; // (row, column) pattern of a lane's slots is the same in every batch
template <typename OutT>
__global__ void __launch_bounds__(kStreamThreads)
fe_stream_kernel(const int64_t N, const int W, const OutT *__restrict__ logret, OutT *__restrict__ obs) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * kStreamThreads + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * kStreamThreads) >> 5;
    const int n = W * 5;
    // slot u of this lane is output element g = lane + 32u of a batch: input element g - g/5, or the position feature
    int off[kStreamSlots];
#pragma unroll
    for (int u = 0; u < kStreamSlots; ++u) {
        const int g = lane + 32 * u;
        off[u] = g % 5 == 4 ? -1 : g - g / 5;
    }
    const int nfull = n / kStreamSpan, tail = n - nfull * kStreamSpan; // elements of the last, partial batch
    auto header = [&](int64_t e, int64_t &row0, OutT &pf) {
        const OutT *slice = obs + (size_t)e * n;
        if constexpr (sizeof(OutT) == 8) {
            row0 = reinterpret_cast<const int64_t *>(slice)[0];
        } else {
            const uint32_t lo = reinterpret_cast<const uint32_t *>(slice)[0], hi = reinterpret_cast<const uint32_t *>(slice)[1];
            row0 = (int64_t)(((uint64_t)hi << 32) | lo);
        }
        pf = slice[4];
    };
    int64_t row0 = 0, row0_next = 0;
    OutT pf = 0, pf_next = 0;
    if (warp0 < N) header(warp0, row0, pf);
    for (int64_t e = warp0; e < N; e += nwarps) {
        // the next env's header travels while this env's window does (a warp's envs are nwarps apart: other slices)
        if (e + nwarps < N) header(e + nwarps, row0_next, pf_next);
        OutT *out = obs + (size_t)e * n + lane;
        const OutT *src = logret + row0 * 4;
        __syncwarp(); // every lane holds this env's header before any lane overwrites it
        for (int b = 0; b < nfull; ++b, out += kStreamSpan, src += kStreamSpanIn) {
            OutT v[kStreamSlots];
#pragma unroll
            for (int u = 0; u < kStreamSlots; ++u) v[u] = off[u] >= 0 ? __ldg(src + off[u]) : pf;
#pragma unroll
            for (int u = 0; u < kStreamSlots; ++u) out[32 * u] = v[u];
        }
        if (tail > 0) {
            OutT v[kStreamSlots];
#pragma unroll
            for (int u = 0; u < kStreamSlots; ++u) v[u] = (off[u] >= 0 && lane + 32 * u < tail) ? __ldg(src + off[u]) : pf;
#pragma unroll
            for (int u = 0; u < kStreamSlots; ++u)
                if (lane + 32 * u < tail) out[32 * u] = v[u];
        }
        row0 = row0_next;
        pf = pf_next;
    }
}

// ------------------------------------------------------------------------------------------
// rows variant (explored alternative, not chosen by `auto`): thread-per-env bookkeeping and warp-autonomous streaming
// with 16-byte global accesses in ONE launch, no bulk copies, no mbarriers.  A block of 256 threads runs env_step() for
// 256 envs (headers {row0, position feature} in shared memory), then its 8 warps stream those envs' windows: a WARP
// owns a group of G consecutive envs (G a power of two >= 4, so a group's slice of the (N, W, 5) tensor starts and ends
// on a 16-byte boundary for any W) and walks its G*W window rows in chunks of 5120 output bytes: lane l loads rows l,
// l+32, ... of the chunk (one 16-byte load per f32 row; the env of a row comes from a multiply-high by ceil(2^32 / W),
// its row0 / position feature by shuffle from the lane holding that env's header), writes them 4 -> 5 interleaved
// into the warp's PRIVATE 5 KB staging buffer (stride-5 words: conflict-free), and after a __syncwarp the warp copies
// the buffer out with 16-byte shared loads and 512-contiguous-byte global stores.
// Measured (profiles/r01_v6_rows_*.txt; c2, 1 Mi envs, W = 60): 0.302 ms (pipe 0.272, tile 0.36, split 0.52) at 3
// resident blocks per SM (4 blocks / 64 registers: 0.317, spills; 2 blocks: 0.330).  ncu: 81 % of the stall samples sit
// in the streaming part, on the window loads' scoreboard and on the STG.128 — 24 warps per SM, each a serial
// load -> stage -> store chain, do not keep enough bytes in flight; c4 (series in HBM): 0.54 ms vs pipe 0.376.
// ------------------------------------------------------------------------------------------
constexpr int kRowsThreads = 256;
constexpr int kRowsWarps = kRowsThreads / 32;
#ifndef FE_ROWS_MINB
#define FE_ROWS_MINB 3 /* resident blocks per SM the rows kernel is compiled for (85 registers: no spills) */
#endif
constexpr int kRowsChunkBytes = 5120; // staging per warp: 256 f32 rows or 128 f64 rows of 5 values
template <typename OutT> struct RowsChunk {
    static constexpr int rows = kRowsChunkBytes / (5 * (int)sizeof(OutT));
    static constexpr int U = rows / 32; // rows per lane and chunk
};
__device__ __forceinline__ int64_t shfl_i64(int64_t v, int src) {
    const int lo = __shfl_sync(0xFFFFFFFFu, (int)(uint32_t)((uint64_t)v & 0xFFFFFFFFu), src);
    const int hi = __shfl_sync(0xFFFFFFFFu, (int)(uint32_t)((uint64_t)v >> 32), src);
    return (int64_t)(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo);
}
// windows of the g <= 32 envs whose headers sit in lanes 0 .. g-1 -> dst (16-byte aligned), through `stage`
template <typename OutT>
__device__ __forceinline__ void rows_stream_group(const Row4<OutT> *__restrict__ series_rows, OutT *__restrict__ dst,
                                                  unsigned char *stage, const int W, const uint32_t magicW, const int g,
                                                  const int64_t row0_l, const OutT pf_l, const int lane) {
    constexpr int CH = RowsChunk<OutT>::rows, U = RowsChunk<OutT>::U;
    const int R = g * W;
    OutT *so = reinterpret_cast<OutT *>(stage);
    for (int base = 0; base < R; base += CH) {
        const int n = min(CH, R - base);
        Row4<OutT> v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int rl = lane + 32 * u;
            const int r = rl < n ? base + rl : base; // lanes past the end shuffle along with a valid row
            const int e = W == 1 ? r : (int)__umulhi((unsigned)r, magicW);
            const int64_t r0 = shfl_i64(row0_l, e);
            if (rl < n) v[u] = ldg_row(series_rows + r0 + (r - e * W));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int rl = lane + 32 * u;
            const int r = rl < n ? base + rl : base;
            const OutT pfe = __shfl_sync(0xFFFFFFFFu, pf_l, W == 1 ? r : (int)__umulhi((unsigned)r, magicW));
            if (rl < n) {
                OutT *o = so + rl * 5;
                if constexpr (sizeof(OutT) == 4) {
                    o[0] = v[u].v.x; o[1] = v[u].v.y; o[2] = v[u].v.z; o[3] = v[u].v.w;
                } else {
                    o[0] = v[u].a.x; o[1] = v[u].a.y; o[2] = v[u].b.x; o[3] = v[u].b.y;
                }
                o[4] = pfe;
            }
        }
        __syncwarp();
        OutT *d = dst + (size_t)base * 5;
        const int nel = n * 5, n16 = (nel * (int)sizeof(OutT)) >> 4;
        uint4 *d16 = reinterpret_cast<uint4 *>(d);
        const uint4 *s16 = reinterpret_cast<const uint4 *>(stage);
#pragma unroll 5
        for (int i = lane; i < n16; i += 32) d16[i] = s16[i];
        // only a ragged last group (N % G envs) can end off a 16-byte boundary
        for (int f = n16 * (16 / (int)sizeof(OutT)) + lane; f < nel; f += 32) d[f] = so[f];
        __syncwarp();
    }
}

template <typename OutT, bool kObserve>
__global__ void __launch_bounds__(kRowsThreads, FE_ROWS_MINB)
fe_rows_kernel(const FeParams p, const FeSeries s, const FeState st, const Consts k, const float *__restrict__ actions,
               OutT *__restrict__ obs, OutT *__restrict__ rewards, int32_t *__restrict__ dones, FeStats *stats,
               const uint64_t step_arg, const uint64_t *__restrict__ step_dev, const int G, const uint32_t magicW) {
    __shared__ __align__(16) unsigned char stage_all[kRowsWarps][kRowsChunkBytes];
    __shared__ int64_t sh_row0[kRowsThreads];
    __shared__ OutT sh_pf[kRowsThreads];
    const uint64_t step = step_dev ? *step_dev : step_arg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t env0 = (int64_t)blockIdx.x * kRowsThreads;
    const int nvalid = (int)min((int64_t)kRowsThreads, p.num_envs - env0);
    const bool active = tid < nvalid;
    EnvResult r;
    r.done = 0; r.newly_terminated = 0; r.fin_return = 0.0; r.fin_len = 0; r.row0 = 0; r.posfeat = 0.0;
    if (active) {
        if (kObserve) r = env_observe(p, s, st, k, env0 + tid);
        else r = env_step<OutT>(p, s, st, k, env0 + tid, actions, rewards, dones, stats != nullptr, step);
    }
    sh_row0[tid] = r.row0;
    sh_pf[tid] = (OutT)r.posfeat;
    if (!kObserve) accumulate_stats(stats, r, active);
    __syncthreads();
    const size_t n = (size_t)p.window * 5;
    for (int e = warp * G; e < nvalid; e += kRowsWarps * G) { // G divides 256: groups never straddle blocks
        const int g = min(G, nvalid - e);
        const int64_t row0 = lane < g ? sh_row0[e + lane] : 0;
        const OutT pf = lane < g ? sh_pf[e + lane] : (OutT)0;
        rows_stream_group<OutT>(reinterpret_cast<const Row4<OutT> *>(s.logret), obs + (size_t)(env0 + e) * n, stage_all[warp],
                                p.window, magicW, g, row0, pf, lane);
    }
}

// ------------------------------------------------------------------------------------------
// portfolio variant (A > 1 assets, one cash account): EXTENSION, the reference is single-asset (:223).
// Semantics: the reference's phases in the reference's order; inside a phase the
// cash moves ONCE: by the f64 butterfly sum of the per-asset deltas where these do not depend on cash (sales,
// covers, margin re-mark / calls / release; bankruptcy is tested on the result), and by a greedy walk over
// the assets in index order with an f64 running cash where they do (long and short entries: an entry is
// legal if the cash left by the assets before it pays for it); A = 1 is the reference exactly.  One WARP per
// env does the bookkeeping with lane = asset (fe_portfolio_book_kernel): every per-asset quantity is
// lane-local, only the f32 cash chain is serial (all lanes keep an identical copy of cash); per-env sums
// (reward, share count) are xor-butterfly warp reductions.  fe_portfolio_stream_kernel then streams the
// (W, A, 4) window in chunks, one block per env: bulk copy in -> interleave 4->5 -> bulk store, double buffered.
// ------------------------------------------------------------------------------------------
constexpr int kPortThreads = 128;
constexpr unsigned kFull = 0xFFFFFFFFu;

__device__ __forceinline__ double warp_sum64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = dadd(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum32(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fadd(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// executed by all 32 lanes of the env's warp; lane >= A is a neutral asset (no shares, no action, price 1).
// Cash-independent phases are butterfly sums; for the two greedy entry walks the per-asset costs are exchanged
// through two 32-double scratch rows in shared memory (broadcast LDS, whose addresses do not depend on the cash
// chain, so the loads run ahead of it) and the serial part is one DADD + compare per asset.  (A first version walked
// every phase with __shfl_sync inside the loops: ncu showed ~10 k instructions per env in this warp,
// WARPSYNC.COLLECTIVE wrappers around every shuffle.)
template <typename OutT>
__device__ __forceinline__ void portfolio_step(const FeParams &p, const FeSeries &s, const FeState &st, const Consts &k,
                                               const int64_t i, const int32_t seg_in, const int32_t ptr_in,
                                               const int64_t row0, const float *__restrict__ actions,
                                               OutT *__restrict__ rewards, int32_t *__restrict__ dones, FeStats *stats,
                                               const uint64_t step, OutT *pf_smem, double *scratch) {
    const int A = p.num_assets, W = p.window;
    const int lane = threadIdx.x & 31;
    const bool act = lane < A;
    const int64_t ia = i * A + lane;
    double *xa = scratch, *xb = scratch + 32; // two 32-double rows; __syncwarp() orders the exchanges
    double O = 1.0, H = 1.0, L = 1.0, C = 1.0, margin = 0.0;
    float lng = 0.0f, sht = 0.0f, d = 0.0f;
    if (act) {
        const double2 *px = reinterpret_cast<const double2 *>(s.prices + ((row0 + W - 1) * A + lane) * 4);
        const double2 oh = __ldg(px), lc = __ldg(px + 1);
        O = oh.x; H = oh.y; L = lc.x; C = lc.y;
        lng = st.long_sh[ia]; sht = st.short_sh[ia]; margin = st.margin[ia];
        d = rintf(fmul(__ldg(actions + ia), k.scale));                       // :298-302
        d = d < -k.ms ? -k.ms : (d > k.ms ? k.ms : d);
    }
    float cash = st.cash[i];
    const int32_t len = __ldg(s.seg_len + seg_in);
    float comm = 0.0f;
    float pos = d < 0.0f ? 0.0f : d, neg = d > 0.0f ? 0.0f : d;             // :344-351
    const double Omc = dsub(O, k.c), Opc = dadd(O, k.c);
    { // :353-361 sell longs ; :367-374 cover shorts ; :375-383 re-mark margin.  The three cash deltas of an asset do not
      // depend on cash: each phase adds the butterfly sum of its per-asset deltas to cash and rounds once
        const float nl = relu32(fadd(lng, neg));
        const float sold = fsub(lng, nl);
        neg = fadd(neg, sold);
        comm = fadd(comm, fmul(sold, k.cf));
        lng = nl;
        cash = d2f(dadd((double)cash, warp_sum64(act ? dmul((double)sold, Omc) : 0.0)));
        const float ns = relu32(fsub(sht, pos));
        const float bought = fsub(sht, ns);
        pos = fsub(pos, bought);
        comm = fadd(comm, fmul(bought, k.cf));
        sht = ns;
        cash = d2f(dsub((double)cash, warp_sum64(act ? dmul((double)bought, Opc) : 0.0)));
        const double nm = dmul((double)fmul(k.imrf, sht), O);
        cash = d2f(dsub((double)cash, warp_sum64(act ? dsub(nm, margin) : 0.0)));
        margin = nm;
    }
    { // :385-399 long entries, greedy in asset order
        xa[lane] = dmul((double)pos, Opc);
        __syncwarp();
        unsigned blocked_mask = 0;
        double c = (double)cash;
        for (int a = 0; a < A; ++a) {
            const double ca = xa[a];
            const bool blocked = dsub(c, ca) < 0.0;
            if (!blocked) c = dsub(c, ca);
            blocked_mask |= (blocked ? 1u : 0u) << a;
        }
        cash = d2f(c);
        if ((blocked_mask >> lane) & 1u) pos = 0.0f;
        comm = fadd(comm, fmul(pos, k.cf));
        lng = fadd(lng, pos);
        __syncwarp();
    }
    { // :401-421 short entries, greedy in asset order
        float q = -neg;
        float sc = fmul(q, k.cf);
        double req = dmul(k.imr, dmul((double)q, O));
        xa[lane] = req;
        xb[lane] = (double)sc;
        __syncwarp();
        unsigned blocked_mask = 0;
        double c = (double)cash;
        for (int a = 0; a < A; ++a) {
            const double ra = xa[a], sa = xb[a];
            const bool blocked = dsub(dsub(c, ra), sa) < 0.0;
            if (!blocked) c = dsub(c, dadd(ra, sa));
            blocked_mask |= (blocked ? 1u : 0u) << a;
        }
        cash = d2f(c);
        if ((blocked_mask >> lane) & 1u) { q = -0.0f; sc = fmul(q, k.cf); req = dmul(k.imr, dmul((double)q, O)); }
        comm = fadd(comm, fmul(q, k.cf));
        margin = dadd(margin, req);
        sht = fadd(sht, q);
        __syncwarp();
    }
    // :428-431 position feature, built before rewards / dones / reset
    if (act) pf_smem[lane] = (OutT)__ddiv_rn(dmul((double)fsub(lng, sht), C), k.SB);
    int done = cash < 0.0f;                                                  // :448
    const double mc1 = relu64(dsub(dmul(dmul((double)sht, H), k.mmr1), margin)); // :459-468 at High
    margin = dadd(margin, mc1);
    const double rel = relu64(dsub(margin, dmul(dmul((double)sht, L), k.imr)));  // :470-475 at Low
    margin = dsub(margin, rel);
    const double mc2 = relu64(dsub(dmul(dmul((double)sht, C), k.mmr1), margin)); // :451 at Close
    margin = dadd(margin, mc2);
    cash = d2f(dsub((double)cash, warp_sum64(act ? mc1 : 0.0)));                 // all margin calls at High, then :465
    done |= cash < 0.0f;
    cash = d2f(dadd((double)cash, warp_sum64(act ? rel : 0.0)));
    cash = d2f(dsub((double)cash, warp_sum64(act ? mc2 : 0.0)));
    done |= cash < 0.0f;
    double r = dadd(-mc1, -mc2);
    if (done) { lng = 0.0f; sht = 0.0f; }                                    // :452-453
    r = dadd(r, dmul((double)fsub(lng, sht), dsub(C, O)));                   // :454-455
    r = dsub(r, (double)comm);                                               // :456
    double rew = warp_sum64(act ? r : 0.0);
    const float nsh = warp_sum32(act ? fadd(sht, lng) : 0.0f);               // :288
    int32_t seg = seg_in, ptr = ptr_in;
    done |= (ptr + W >= len);                                                // :477-496
    rew = dsub(rew, (double)fmul(fmul(done ? 1.0f : 0.0f, nsh), k.cf));      // :288-289
    if (done) {                                                              // :498-521
        cash = k.SBf; margin = 0.0; lng = 0.0f; sht = 0.0f; ptr = 0;
        const int64_t gid = p.env_id_base + i;
        if (p.reset_mode == FE_RESET_ALL || (p.reset_mode == FE_RESET_LAST && gid == p.total_envs - 1))
            draw_segment(p, s, gid, step, 0u, seg, ptr);
    }
    if (act) { st.long_sh[ia] = lng; st.short_sh[ia] = sht; st.margin[ia] = margin; }
    if (lane == 0) {
        if (p.evaluate) {                                                    // :523-536
            const int was = st.terminated[i];
            if (was) rew = 0.0;
            if (done && !was) { st.terminated[i] = 1; if (stats) atomicAdd(&stats->n_terminated, 1ULL); }
            st.ep_return[i] = d2f(dadd((double)st.ep_return[i], rew));
            if (done && stats) atomicAdd(&stats->n_done, 1ULL);
        } else if (stats) {
            const float er = d2f(dadd((double)st.ep_return[i], rew));
            const int32_t el = st.ep_len[i] + 1;
            if (done) {
                atomicAdd(&stats->n_done, 1ULL);
                atomicAdd(&stats->sum_len, (unsigned long long)el);
                atomicAdd(&stats->sum_return, (double)er);
                atomicAdd(&stats->sum_return_sq, (double)er * (double)er);
            }
            st.ep_return[i] = done ? 0.0f : er;
            st.ep_len[i] = done ? 0 : el;
        }
        st.seg[i] = seg; st.ptr[i] = ptr; st.cash[i] = cash;
        rewards[i] = (OutT)rew;
        dones[i] = done;
        if (k.dones_mirror) { reinterpret_cast<OutT *>(k.rewards_mirror)[i] = (OutT)rew; k.dones_mirror[i] = done; }
    }
}

// stream kernel smem: [2 mbarriers 16 B][pf 32 x OutT -> 256 B][unused 768 B][in[2]: CH*4 OutT each][out[2]: CH*5 OutT each]
constexpr int kPortHeader = 16 + 256 + 768;
template <typename OutT> __host__ __device__ inline size_t port_smem_bytes(int CH) {
    return kPortHeader + (size_t)2 * CH * 9 * sizeof(OutT);
}

// ------------------------------------------------------------------------------------------
// portfolio kernels (A > 1): bookkeeping and streaming as two launches.  A first, fused form (one block per env: warp 0
// bookkeeping, then the whole block streaming; profiles/r01_v4_portfolio_fused_ncu.txt) ran at 0.82 of the HBM
// roofline: warp 0's serial cash chain (8 passes over the A assets, each a cvt -> DADD -> cvt dependency) holds its
// block's streaming back for several microseconds per env (11.9 of 21.6 stall cycles per issue were barrier waits).  Here fe_portfolio_book_kernel runs the bookkeeping with one WARP per env at full occupancy
// (~0.05 ms for 65 536 envs) and leaves each env's {row0, position features} as a header INSIDE the env's slice of the
// observation tensor (the position features already at their final place in window row 0); fe_portfolio_stream_kernel
// then reads the header and streams the window exactly as before, with nothing serial in front of it.
// ------------------------------------------------------------------------------------------
constexpr int kBookWarps = 4;

template <typename OutT, bool kObserve>
__global__ void __launch_bounds__(kBookWarps * 32)
fe_portfolio_book_kernel(const FeParams p, const FeSeries s, const FeState st, const Consts k,
                         const float *__restrict__ actions, OutT *__restrict__ obs, OutT *__restrict__ rewards,
                         int32_t *__restrict__ dones, FeStats *stats, const uint64_t step_arg,
                         const uint64_t *__restrict__ step_dev) {
    __shared__ double scratch[kBookWarps][64];
    __shared__ OutT pf_s[kBookWarps][32];
    const uint64_t step = step_dev ? *step_dev : step_arg;
    const int W = p.window, A = p.num_assets;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * kBookWarps + warp;
    if (i >= p.num_envs) return; // whole warps leave together; no block-wide barrier below
    const int32_t seg = st.seg[i];
    const int32_t ptr = st.ptr[i] + (kObserve ? 0 : 1);
    const int64_t row0 = __ldg(s.seg_start + seg) + ptr;
    OutT *pf = pf_s[warp];
    if (kObserve) {
        if (lane < A) {
            const double C = __ldg(s.prices + ((row0 + W - 1) * A + lane) * 4 + 3);
            const float net = fsub(st.long_sh[i * A + lane], st.short_sh[i * A + lane]);
            pf[lane] = (OutT)__ddiv_rn(dmul((double)net, C), k.SB);
        }
    } else {
        portfolio_step<OutT>(p, s, st, k, i, seg, ptr, row0, actions, rewards, dones, stats, step, pf, scratch[warp]);
    }
    __syncwarp();
    // header: row0 in the first 8 bytes of the env's slice (two words for f32 obs: the slice is only 4-byte aligned
    // when W*A is odd), position feature of asset a at its final place, element (row 0, asset a, column 4)
    OutT *slice = obs + (size_t)i * W * A * 5;
    if (lane == 0) {
        if constexpr (sizeof(OutT) == 8) {
            reinterpret_cast<int64_t *>(slice)[0] = row0;
        } else {
            reinterpret_cast<uint32_t *>(slice)[0] = (uint32_t)((uint64_t)row0 & 0xFFFFFFFFu);
            reinterpret_cast<uint32_t *>(slice)[1] = (uint32_t)((uint64_t)row0 >> 32);
        }
    }
    if (lane < A) slice[(size_t)lane * 5 + 4] = pf[lane];
}

template <typename OutT>
__global__ void __launch_bounds__(kPortThreads)
fe_portfolio_stream_kernel(const FeParams p, const FeSeries s, OutT *__restrict__ obs, const int CH) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int W = p.window, A = p.num_assets;
    const int tid = threadIdx.x;
    const int64_t i = blockIdx.x;
    const int P = W * A;                         // (row, asset) pairs of one env's window
    const int nchunks = (P + CH - 1) / CH;
    OutT *pf = reinterpret_cast<OutT *>(smem + 16);
    unsigned char *in0 = smem + kPortHeader;
    const size_t in_bytes = (size_t)CH * 4 * sizeof(OutT), out_bytes = (size_t)CH * 5 * sizeof(OutT);
    unsigned char *out0 = in0 + 2 * in_bytes;
    const uint32_t bar0 = smem_u32(smem);
    OutT *dst = obs + (size_t)i * P * 5;
    // the header left by fe_portfolio_book_kernel; every thread reads row0 before anything overwrites it
    int64_t row0;
    if constexpr (sizeof(OutT) == 8) {
        row0 = reinterpret_cast<const int64_t *>(dst)[0];
    } else {
        const uint32_t lo = reinterpret_cast<const uint32_t *>(dst)[0], hi = reinterpret_cast<const uint32_t *>(dst)[1];
        row0 = (int64_t)(((uint64_t)hi << 32) | lo);
    }
    if (tid < A) pf[tid] = dst[(size_t)tid * 5 + 4];
    const unsigned char *src = reinterpret_cast<const unsigned char *>(s.logret) + (size_t)row0 * A * 4 * sizeof(OutT);
    const bool bulk_out = (P & 3) == 0;          // then every chunk of every env is 16-byte aligned and sized
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        mbar_fence_init();
    }
    __syncthreads(); // header consumed, pf visible, barriers initialised
    auto load_chunk = [&](int c) { // thread 0 only
        const int n = min(CH, P - c * CH);
        const uint32_t bytes = (uint32_t)n * 4 * sizeof(OutT);
        const uint32_t bar = bar0 + 8 * (c & 1);
        mbar_arrive_expect_tx(bar, bytes);
        bulk_load(smem_u32(in0 + (c & 1) * in_bytes), src + (size_t)c * in_bytes, bytes, bar);
    };
    if (tid == 0) {
        load_chunk(0);
        if (nchunks > 1) load_chunk(1);
    }
    const float invA = 1.0f / (float)A;
    for (int c = 0; c < nchunks; ++c) {
        const int sidx = c & 1;
        const int n = min(CH, P - c * CH);
        mbar_wait(bar0 + 8 * sidx, (c >> 1) & 1);
        if (c >= 2) { // out[sidx] is being read by the store of chunk c-2: allow only chunk c-1's store in flight
            if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncthreads();
        }
        const Row4<OutT> *in_rows = reinterpret_cast<const Row4<OutT> *>(in0 + sidx * in_bytes);
        OutT *out_tile = reinterpret_cast<OutT *>(out0 + sidx * out_bytes);
        const int g0 = c * CH;
        for (int r = tid; r < n; r += kPortThreads) {
            const int g = g0 + r;
            const int a = g - __float2int_rz(((float)g + 0.5f) * invA) * A;
            const Row4<OutT> v = in_rows[r];
            OutT *o = out_tile + (size_t)r * 5;
            if constexpr (sizeof(OutT) == 4) {
                o[0] = v.v.x; o[1] = v.v.y; o[2] = v.v.z; o[3] = v.v.w;
            } else {
                o[0] = v.a.x; o[1] = v.a.y; o[2] = v.b.x; o[3] = v.b.y;
            }
            o[4] = pf[a];
        }
        OutT *d = dst + (size_t)g0 * 5;
        if (bulk_out) {
            fence_proxy_async_smem();
            __syncthreads(); // out tile complete, in tile fully consumed
            if (tid == 0) {
                bulk_store(d, smem_u32(out_tile), (uint32_t)((size_t)n * 5 * sizeof(OutT)));
                bulk_commit();
                if (c + 2 < nchunks) load_chunk(c + 2);
            }
        } else {
            __syncthreads();
            for (int f = tid; f < n * 5; f += kPortThreads) d[f] = out_tile[f];
            __syncthreads();
            if (tid == 0 && c + 2 < nchunks) load_chunk(c + 2);
        }
    }
    if (tid == 0) bulk_wait_read_all();
}

// ------------------------------------------------------------------------------------------
// reset_all, log-returns, effective segment length
// ------------------------------------------------------------------------------------------
__global__ void fe_reset_all_kernel(const FeParams p, const FeSeries s, const FeState st, const uint64_t step,
                                    const int redraw) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.num_envs) return;
    st.cash[i] = (float)p.starting_balance;
    for (int a = 0; a < p.num_assets; ++a) {
        st.margin[i * p.num_assets + a] = 0.0;
        st.long_sh[i * p.num_assets + a] = 0.0f;
        st.short_sh[i * p.num_assets + a] = 0.0f;
    }
    int32_t ptr = 0;
    if (redraw) {
        int32_t seg;
        draw_segment(p, s, p.env_id_base + i, step, 1u, seg, ptr);
        st.seg[i] = seg;
    }
    st.ptr[i] = ptr;
    if (st.terminated) st.terminated[i] = 0;
    if (st.ep_return) st.ep_return[i] = 0.0f;
    if (st.ep_len) st.ep_len[i] = 0;
}

// :179-194.  log() here is CUDA's double-precision log (<= 1 ulp), the reference's is torch's.
__global__ void fe_log_returns_kernel(const double *__restrict__ prices, const int64_t num_rows, const int A,
                                      double *__restrict__ lr64, float *__restrict__ lr32) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; // (t, a)
    if (idx >= num_rows * A) return;
    const int64_t t = idx / A;
    const double *px = prices + idx * 4;
    const double o = px[0];
    const double prev_close = t == 0 ? o : prices[(idx - A) * 4 + 3];
    double v[4];
    v[0] = dmul(100.0, log(__ddiv_rn(o, prev_close)));
#pragma unroll
    for (int c = 1; c < 4; ++c) v[c] = dmul(100.0, log(__ddiv_rn(px[c], o)));
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (lr64) lr64[idx * 4 + c] = v[c];
        if (lr32) lr32[idx * 4 + c] = (float)v[c];
    }
}

// :486-496 folded into the table: one warp per segment scans for the first NaN in column 0
__global__ void fe_effective_len_kernel(const double *__restrict__ lr64, const int64_t *__restrict__ seg_start,
                                        const int32_t *__restrict__ raw_len, const int D, const int W, const int A,
                                        int32_t *__restrict__ seg_len) {
    const int d = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (d >= D) return;
    const int64_t start = seg_start[d];
    const int n = raw_len[d];
    int best = n;
    for (int kk = W + 1 + lane; kk < n; kk += 32) {
        const double v = lr64[(size_t)(start + kk) * A * 4];
        if (v != v) { best = kk; break; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
    if (lane == 0) seg_len[d] = best;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
int check_common(const FeParams *p, const FeSeries *s, const FeState *st) {
    if (!p || !s || !st) return FE_EINVAL;
    if (p->num_envs <= 0 || p->window <= 0 || p->num_segments <= 0 || p->num_rows <= 0) return FE_EINVAL;
    if (p->num_assets < 1 || p->num_assets > 32) return FE_EINVAL;
    if (!s->prices || !s->logret || !s->seg_start || !s->seg_len) return FE_EINVAL;
    if (!st->seg || !st->ptr || !st->cash || !st->long_sh || !st->short_sh || !st->margin) return FE_EINVAL;
    if (p->evaluate && (!st->terminated || !st->ep_return)) return FE_EINVAL;
    if (((uintptr_t)s->prices | (uintptr_t)s->logret) & 15) return FE_EALIGN;
    return 0;
}

// Tile shape.  Measured on B200 (tools/sweep_tile.sh, 1 Mi envs, W=60): the step is latency-bound per
// block (state loads -> table lookup -> window copy -> interleave -> store), so MANY SMALL blocks in
// flight beat few large ones: 4 envs x 32 threads (23 blocks/SM) runs 0.35 ms where 32 envs x 128
// threads (3 blocks/SM) runs 0.46 ms.  E is kept a multiple of 4 so that every full tile's byte count
// (E*W*20) is a multiple of 16 for any W (bulk-copy granularity).
int pick_tile_envs(int W, bool f64, int *threads_out = nullptr) {
    const size_t sz = f64 ? 8 : 4;
    const size_t per_env = (size_t)W * 9 * sz + sz;
    long e_max = (long)((kSmemMax - kSmemHeader - 16) / per_env) & ~3L;
    if (e_max < 4) return 0;
    long e = 4;
    while (e * W < 192 && e < kMaxTileEnvs) e += 4; // tiny windows: keep >= ~200 rows per block
    if (e > e_max) e = e_max;
    long rows = e * W;
    int threads = (int)(((rows / 8) + 31) & ~31L);
    if (threads < 32) threads = 32;
    if (threads > kThreads) threads = kThreads;
    while (threads < e) threads += 32;
    if (threads_out) *threads_out = threads;
    return (int)e;
}

// tuning overrides (sweeps only)
int env_override(const char *name) {
    const char *v = getenv(name);
    return v ? atoi(v) : 0;
}

// scatter variant: envs per tile (<= 32, multiple of 4), ring stages S (4..8) and fill depth D = S - 3
// (D + 1 tiles being filled / landing, up to 2 being stored); false = window too large
bool pick_scatter(int W, bool f64, int *TE, int *S, int *D) {
    static const int ov_te = env_override("FE_SC_TE"), ov_s = env_override("FE_SC_STAGES"), ov_d = env_override("FE_SC_DEPTH");
    int te = 32;
    if (ov_te >= 4) te = ov_te & ~3;
    if (te > 32) te = 32;
    auto bytes = [&](int t, int st) { return f64 ? scatter_smem_bytes<double>(t, W, st) : scatter_smem_bytes<float>(t, W, st); };
    while (te >= 4 && bytes(te, 4) > (size_t)kSmemMax) te -= 4;
    if (te < 4) return false;
    int st = kScMaxStages;
    while (st > 4 && bytes(te, st) > (size_t)kSmemMax) --st;
    if (ov_s >= 4 && ov_s <= st) st = ov_s;
    int d = st - 3;
    if (ov_d >= 1 && ov_d <= st - 2) d = ov_d;
    if (d > 6) d = 6;
    *TE = te; *S = st; *D = d;
    return true;
}

// envs per tile of the pipe variant: up to 32 (one bookkeeper lane each), a multiple of 4 (16-byte store
// granularity), with TE*W rows fitting the movers' register staging; 0 = window too large, use the tile variant
int pick_pipe_envs(int W, bool f64, int sin) {
    const int max_rows = kMovers * (f64 ? PipeRows<double>::value : PipeRows<float>::value);
    int te = (max_rows / W) & ~3;
    if (te > 32) te = 32;
    while (te >= 4 && (f64 ? pipe_smem_bytes<double>(te, W, sin) : pipe_smem_bytes<float>(te, W, sin)) > (size_t)kSmemMax) te -= 4;
    return te < 4 ? 0 : te;
}
// in-ring stages: 0 ("cached" flavour) while the log-return table is comfortably L2-resident (126 MB L2, shared
// with the observation stream passing through it), else the "stream" flavour
int pick_pipe_stages(const FeParams &p, bool f64) {
    const size_t table = (size_t)p.num_rows * p.num_assets * 4 * (f64 ? 8 : 4);
    return table <= ((size_t)48 << 20) ? 0 : kPipeSInStream;
}


// Which kernel a launch uses.  `auto`: portfolio when A > 1; split for windows too long for the pipe variant; the
// persistent pipe variant for populations that give every
// SM a few tiles AND windows of >= 24 rows (measured on c2, tools/window_sweep.sh -> profiles/r01_v4_window_sweep.txt:
// W = 4 / 16: tile 0.19 / 0.24 ms vs pipe 0.73 / 0.37 ms — with so few rows per env the 6 bookkeeper warps are the
// bottleneck, while the tile variant gives every env its own thread; W = 60 / 128 / 390: pipe 0.27 / 0.27 / 0.23 vs tile
// 0.36 / 0.38 / 0.29); else tile while the window fits in shared memory; else direct.
enum StepKernel { K_PORTFOLIO, K_SPLIT, K_ROWS, K_SCATTER, K_PIPE, K_TILE, K_DIRECT, K_ERR_SMEM };
struct StepChoice {
    StepKernel kern;
    int te;      // envs per tile (pipe / scatter / tile)
    int sin;     // pipe: in-ring stages (0 = cached flavour)
    int S, D;    // scatter: ring stages, fill depth
    int threads; // tile: threads per block
};
StepChoice choose_kernel(const FeParams &p, bool f64) {
    StepChoice c = {K_DIRECT, 0, 0, 0, 0, kThreads};
    if (p.num_assets > 1 || p.variant == FE_VARIANT_PORTFOLIO) { c.kern = K_PORTFOLIO; return c; }
    static const int auto_split = env_override("FE_AUTO_SPLIT");     // sweeps: 1 = "auto" prefers split, -1 = never
    if ((p.variant == FE_VARIANT_SPLIT || (p.variant == FE_VARIANT_AUTO && auto_split > 0))) { c.kern = K_SPLIT; return c; }
    // rows: the multiply-high row -> env map is exact while 32 * W * W < 2^32
    if (p.variant == FE_VARIANT_ROWS) { c.kern = p.window <= 8192 ? K_ROWS : K_ERR_SMEM; return c; }
    static const int auto_scatter = env_override("FE_AUTO_SCATTER"); // sweeps: 1 = "auto" prefers scatter
    static const int no_pipe = env_override("FE_NO_PIPE");           // sweeps: "auto" never picks pipe
    // small populations: a persistent grid needs a few tiles per SM to hide its prologue
    const bool worth_persistent = p.num_envs >= (int64_t)4 * 148 * 32;
    if (p.variant == FE_VARIANT_SCATTER || (p.variant == FE_VARIANT_AUTO && worth_persistent && auto_scatter > 0)) {
        if (pick_scatter(p.window, f64, &c.te, &c.S, &c.D)) { c.kern = K_SCATTER; return c; }
        if (p.variant == FE_VARIANT_SCATTER) { c.kern = K_ERR_SMEM; return c; }
    }
    if (p.variant == FE_VARIANT_PIPE || (p.variant == FE_VARIANT_AUTO && !no_pipe)) {
        static const int ov_sin = env_override("FE_PIPE_FLAVOUR"); // sweeps: 1 = cached, 2 = stream
        c.sin = ov_sin == 1 ? 0 : ov_sin == 2 ? kPipeSInStream : pick_pipe_stages(p, f64);
        c.te = pick_pipe_envs(p.window, f64, c.sin);
        if (c.te == 0 && p.variant == FE_VARIANT_PIPE) { c.kern = K_ERR_SMEM; return c; }
        if (c.te > 0 && (p.variant == FE_VARIANT_PIPE || (worth_persistent && p.window >= 24))) { c.kern = K_PIPE; return c; }
        // windows too long for the pipe variant's rings (> 512 rows f32, > 256 f64): the split variant's warp-per-env
        // streaming (W = 1024, 64 Ki envs: 0.26 ms vs 0.66 ms for tile)
        if (c.te == 0 && p.variant == FE_VARIANT_AUTO && auto_split >= 0) { c.kern = K_SPLIT; return c; }
    }
    c.te = p.variant == FE_VARIANT_DIRECT ? 0 : pick_tile_envs(p.window, f64, &c.threads);
    if (p.variant == FE_VARIANT_TILE && c.te == 0) { c.kern = K_ERR_SMEM; return c; }
    if (c.te > 0) {
        static const int ov_e = env_override("FE_TILE_ENVS"), ov_t = env_override("FE_TILE_THREADS");
        const size_t sz = f64 ? 8 : 4;
        if (ov_e >= 4 && kSmemHeader + 16 + (size_t)(ov_e & ~3) * (p.window * 9 * sz + sz) <= (size_t)kSmemMax) c.te = ov_e & ~3;
        if (ov_t >= 32 && ov_t <= kThreads) c.threads = ov_t & ~31;
        if (c.te > c.threads) c.te = c.threads;
        c.kern = K_TILE;
    }
    return c;
}

int device_sm_count(int device, int *out) {
    static int num_sms[16] = {0};
    const int dev = device & 15;
    if (!num_sms[dev]) {
        cudaError_t e = cudaDeviceGetAttribute(&num_sms[dev], cudaDevAttrMultiProcessorCount, device);
        if (e != cudaSuccess) return (int)e;
    }
    *out = num_sms[dev];
    return 0;
}

template <typename OutT, bool kObserve>
int launch(const FeParams &p, const FeSeries &s, const FeState &st, const float *actions, void *obs, void *rewards,
           int32_t *dones, FeStats *stats, uint64_t step, cudaStream_t stream, const uint64_t *step_dev = nullptr,
           void *rewards_mirror = nullptr, int32_t *dones_mirror = nullptr) {
    Consts k = make_consts(p);
    k.rewards_mirror = rewards_mirror;
    k.dones_mirror = dones_mirror;
    const StepChoice c = choose_kernel(p, sizeof(OutT) == 8);
    const int dev = p.device & 15;
    int sms = 0;
    switch (c.kern) {
    case K_ERR_SMEM: return FE_ESMEM;
    case K_PORTFOLIO: {
        if ((uintptr_t)obs & 15) return FE_EALIGN;
        const int P = p.window * p.num_assets;
        int CH = 512; // pairs per chunk (8 KB in + 10 KB out, x2 stages); C3: 128 -> 1.492 ms, 256 -> 1.480, 512 -> 1.472, 1024 -> 1.471
        static const int ov_ch = env_override("FE_PORT_CHUNK");
        if (ov_ch >= 4) CH = ov_ch & ~3;
        if (CH > ((P + 3) & ~3)) CH = (P + 3) & ~3;
        const size_t smem = port_smem_bytes<OutT>(CH);
        if (smem > (size_t)kSmemMax) return FE_ESMEM;
        auto skern = fe_portfolio_stream_kernel<OutT>;
        static bool sconfigured[16] = {false};
        if (!sconfigured[dev]) {
            cudaError_t e = cudaFuncSetAttribute(skern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
            if (e != cudaSuccess) return (int)e;
            sconfigured[dev] = true;
        }
        fe_portfolio_book_kernel<OutT, kObserve><<<(unsigned)((p.num_envs + kBookWarps - 1) / kBookWarps), kBookWarps * 32, 0, stream>>>(
            p, s, st, k, actions, (OutT *)obs, (OutT *)rewards, dones, stats, step, step_dev);
        skern<<<(unsigned)p.num_envs, kPortThreads, smem, stream>>>(p, s, (OutT *)obs, CH);
        break;
    }
    case K_SPLIT: {
        if ((uintptr_t)obs & 7) return FE_EALIGN;
        int rc = device_sm_count(p.device, &sms);
        if (rc) return rc;
        fe_book_kernel<OutT, kObserve><<<(unsigned)((p.num_envs + kThreads - 1) / kThreads), kThreads, 0, stream>>>(
            p, s, st, k, actions, (OutT *)obs, (OutT *)rewards, dones, stats, step, step_dev);
        static const int ov_b = env_override("FE_STREAM_BLOCKS_PER_SM");
        int64_t blocks = (int64_t)sms * (ov_b > 0 ? ov_b : 6); // 48 warps per SM
        const int64_t need = (p.num_envs * 32 + kStreamThreads - 1) / kStreamThreads;
        if (blocks > need) blocks = need;
        fe_stream_kernel<OutT><<<(unsigned)blocks, kStreamThreads, 0, stream>>>(p.num_envs, p.window, (const OutT *)s.logret, (OutT *)obs);
        break;
    }
    case K_ROWS: {
        if ((uintptr_t)obs & 15) return FE_EALIGN;
        int G = 4; // envs per warp group: a power of two (divides the block's 256 envs), >= 256 rows when it can
        while (G < 32 && G * p.window < 256) G *= 2;
        const uint32_t magicW = p.window == 1 ? 0u : (uint32_t)((((uint64_t)1 << 32) + p.window - 1) / p.window);
        auto kern = fe_rows_kernel<OutT, kObserve>;
        static bool configured[16] = {false};
        if (!configured[dev]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) return (int)e;
            configured[dev] = true;
        }
        kern<<<(unsigned)((p.num_envs + kRowsThreads - 1) / kRowsThreads), kRowsThreads, 0, stream>>>(
            p, s, st, k, actions, (OutT *)obs, (OutT *)rewards, dones, stats, step, step_dev, G, magicW);
        break;
    }
    case K_SCATTER: {
        if ((uintptr_t)obs & 15) return FE_EALIGN;
        auto kern = fe_scatter_kernel<OutT, kObserve>;
        static bool configured[16] = {false};
        if (!configured[dev]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
            if (e != cudaSuccess) return (int)e;
            configured[dev] = true;
        }
        int rc = device_sm_count(p.device, &sms);
        if (rc) return rc;
        const int64_t ntiles = (p.num_envs + c.te - 1) / c.te;
        const unsigned blocks = (unsigned)(ntiles < sms ? ntiles : sms);
        kern<<<blocks, kScThreads, scatter_smem_bytes<OutT>(c.te, p.window, c.S), stream>>>(
            p, s, st, k, actions, (OutT *)obs, (OutT *)rewards, dones, stats, step, step_dev, c.te, c.S, c.D);
        break;
    }
    case K_PIPE: {
        if ((uintptr_t)obs & 15) return FE_EALIGN;
        auto kern = c.sin == 0 ? fe_pipe_kernel<OutT, kObserve, 0> : fe_pipe_kernel<OutT, kObserve, kPipeSInStream>;
        static bool configured[16][2] = {{false}};
        if (!configured[dev][c.sin != 0]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
            if (e != cudaSuccess) return (int)e;
            configured[dev][c.sin != 0] = true;
        }
        int rc = device_sm_count(p.device, &sms);
        if (rc) return rc;
        const int64_t ntiles = (p.num_envs + c.te - 1) / c.te;
        const unsigned blocks = (unsigned)(ntiles < sms ? ntiles : sms);
        kern<<<blocks, kPipeThreads, pipe_smem_bytes<OutT>(c.te, p.window, c.sin), stream>>>(
            p, s, st, k, actions, (OutT *)obs, (OutT *)rewards, dones, stats, step, step_dev, c.te);
        break;
    }
    case K_TILE: {
        if ((uintptr_t)obs & 15) return FE_EALIGN;
        const size_t smem = tile_smem_bytes<OutT>(c.te, p.window);
        auto kern = fe_tile_kernel<OutT, kObserve>;
        static size_t configured[16] = {0}; // per device: largest opt-in already set for this instantiation
        if (smem > configured[dev]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
            if (e != cudaSuccess) return (int)e;
            configured[dev] = kSmemMax;
        }
        const int64_t blocks = (p.num_envs + c.te - 1) / c.te;
        kern<<<(unsigned)blocks, c.threads, smem, stream>>>(p, s, st, k, actions, (OutT *)obs, (OutT *)rewards, dones,
                                                             stats, step, step_dev, c.te);
        break;
    }
    case K_DIRECT: {
        const int64_t blocks = (p.num_envs + kThreads - 1) / kThreads;
        fe_direct_kernel<OutT, kObserve><<<(unsigned)blocks, kThreads, 0, stream>>>(
            p, s, st, k, actions, (OutT *)obs, (OutT *)rewards, dones, stats, step, step_dev);
        break;
    }
    }
    return (int)cudaGetLastError();
}

// step ordinal kept on the device (fe_step_captured): one thread bumps it ahead of the step kernel
__global__ void fe_bump_kernel(uint64_t *counter) { *counter += 1; }

int set_device(int device) {
    int cur = -1;
    cudaError_t e = cudaGetDevice(&cur);
    if (e != cudaSuccess) return (int)e;
    if (cur != device) {
        e = cudaSetDevice(device);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

// side streams of fe_step_host (created once per device, never destroyed: they live as long as the process)
constexpr int kHostStreams = 3;
struct HostPipe {
    cudaStream_t s[kHostStreams];
    cudaEvent_t start, done[kHostStreams];
    bool ready;
};
HostPipe *host_pipe(int device) {
    static HostPipe pipes[16];
    HostPipe *hp = &pipes[device & 15];
    if (!hp->ready) {
        for (int k = 0; k < kHostStreams; ++k) {
            if (cudaStreamCreateWithFlags(&hp->s[k], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&hp->done[k], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        if (cudaEventCreateWithFlags(&hp->start, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        hp->ready = true;
    }
    return hp;
}

} // namespace

extern "C" {

int fe_version(void) { return FE_ABI_VERSION; }

const char *fe_error_string(int code) {
    switch (code) {
    case 0: return "ok";
    case FE_EINVAL: return "finenvs_b200: invalid argument (null pointer, non-positive size or num_assets outside 1..32)";
    case FE_EALIGN: return "finenvs_b200: pointer not 16-byte aligned";
    case FE_ESMEM: return "finenvs_b200: window too large for the tile variant";
    case FE_EIO: return "finenvs_b200: cannot open or map the file";
    case FE_ECSV: return "finenvs_b200: CSV record outside the format of the native reader";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "finenvs_b200: unknown error";
    }
}

int fe_tile_envs(int32_t window, int32_t out_f64, int32_t device) {
    (void)device;
    return window > 0 ? pick_tile_envs(window, out_f64 != 0) : 0;
}

int fe_pipe_envs(int32_t window, int32_t out_f64, int32_t stream_flavour) {
    return window > 0 ? pick_pipe_envs(window, out_f64 != 0, stream_flavour ? kPipeSInStream : 0) : 0;
}

const char *fe_step_kernel_name(const FeParams *p) {
    if (!p) return "";
    const bool f64 = p->out_f64 != 0;
    const StepChoice c = choose_kernel(*p, f64);
    switch (c.kern) {
    case K_PORTFOLIO:
        return f64 ? "fe_portfolio_book_kernel + fe_portfolio_stream_kernel<double>" : "fe_portfolio_book_kernel + fe_portfolio_stream_kernel<float>";
    case K_SPLIT: return f64 ? "fe_book_kernel + fe_stream_kernel<double>" : "fe_book_kernel + fe_stream_kernel<float>";
    case K_ROWS: return f64 ? "fe_rows_kernel<double>" : "fe_rows_kernel<float>";
    case K_SCATTER: return f64 ? "fe_scatter_kernel<double>" : "fe_scatter_kernel<float>";
    case K_PIPE:
        return c.sin == 0 ? (f64 ? "fe_pipe_kernel<double,cached>" : "fe_pipe_kernel<float,cached>")
                          : (f64 ? "fe_pipe_kernel<double,stream>" : "fe_pipe_kernel<float,stream>");
    case K_TILE: return f64 ? "fe_tile_kernel<double>" : "fe_tile_kernel<float>";
    case K_DIRECT: return f64 ? "fe_direct_kernel<double>" : "fe_direct_kernel<float>";
    default: return "none (window too large for the requested variant)";
    }
}

int fe_log_returns(const double *prices_dev, int64_t num_rows, int32_t num_assets, double *logret64_dev,
                   float *logret32_dev, void *stream) {
    if (!prices_dev || num_rows <= 0 || num_assets <= 0) return FE_EINVAL;
    const int64_t n = num_rows * num_assets;
    fe_log_returns_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(prices_dev, num_rows, num_assets,
                                                                                         logret64_dev, logret32_dev);
    return (int)cudaGetLastError();
}

int fe_effective_len(const double *logret64_dev, const int64_t *seg_start_dev, const int32_t *raw_len_dev,
                     int32_t num_segments, int32_t window, int32_t num_assets, int32_t *seg_len_dev, void *stream) {
    if (!logret64_dev || !seg_start_dev || !raw_len_dev || !seg_len_dev || num_segments <= 0) return FE_EINVAL;
    const int64_t threads = (int64_t)num_segments * 32;
    fe_effective_len_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        logret64_dev, seg_start_dev, raw_len_dev, num_segments, window, num_assets, seg_len_dev);
    return (int)cudaGetLastError();
}

int fe_observe(const FeParams *p, const FeSeries *s, const FeState *st, void *obs_dev, void *stream) {
    int rc = check_common(p, s, st);
    if (rc) return rc;
    if (!obs_dev) return FE_EINVAL;
    if ((rc = set_device(p->device))) return rc;
    return p->out_f64 ? launch<double, true>(*p, *s, *st, nullptr, obs_dev, nullptr, nullptr, nullptr, 0, (cudaStream_t)stream)
                      : launch<float, true>(*p, *s, *st, nullptr, obs_dev, nullptr, nullptr, nullptr, 0, (cudaStream_t)stream);
}

int fe_step(const FeParams *p, const FeSeries *s, const FeState *st, const float *actions_dev, void *obs_dev,
            void *rewards_dev, int32_t *dones_dev, FeStats *stats_dev, uint64_t step_counter, void *stream) {
    int rc = check_common(p, s, st);
    if (rc) return rc;
    if (!actions_dev || !obs_dev || !rewards_dev || !dones_dev) return FE_EINVAL;
    if (stats_dev && !p->evaluate && (!st->ep_return || !st->ep_len)) return FE_EINVAL;
    if ((rc = set_device(p->device))) return rc;
    return p->out_f64 ? launch<double, false>(*p, *s, *st, actions_dev, obs_dev, rewards_dev, dones_dev, stats_dev,
                                              step_counter, (cudaStream_t)stream)
                      : launch<float, false>(*p, *s, *st, actions_dev, obs_dev, rewards_dev, dones_dev, stats_dev,
                                             step_counter, (cudaStream_t)stream);
}

int fe_step_captured(const FeParams *p, const FeSeries *s, const FeState *st, const float *actions_dev, void *obs_dev,
                     void *rewards_dev, int32_t *dones_dev, FeStats *stats_dev, uint64_t *step_counter_dev, void *stream) {
    int rc = check_common(p, s, st);
    if (rc) return rc;
    if (!actions_dev || !obs_dev || !rewards_dev || !dones_dev || !step_counter_dev) return FE_EINVAL;
    if (stats_dev && !p->evaluate && (!st->ep_return || !st->ep_len)) return FE_EINVAL;
    if ((rc = set_device(p->device))) return rc;
    fe_bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_counter_dev);
    if ((rc = (int)cudaGetLastError())) return rc;
    return p->out_f64 ? launch<double, false>(*p, *s, *st, actions_dev, obs_dev, rewards_dev, dones_dev, stats_dev, 0,
                                              (cudaStream_t)stream, step_counter_dev)
                      : launch<float, false>(*p, *s, *st, actions_dev, obs_dev, rewards_dev, dones_dev, stats_dev, 0,
                                             (cudaStream_t)stream, step_counter_dev);
}

int fe_observe_lazy(const FeParams *p, const FeSeries *s, const FeState *st, int64_t *obs_row0_dev, void *obs_posfeat_dev,
                    void *stream) {
    int rc = check_common(p, s, st);
    if (rc) return rc;
    if (!obs_row0_dev || !obs_posfeat_dev || p->num_assets != 1) return FE_EINVAL;
    if ((rc = set_device(p->device))) return rc;
    const Consts k = make_consts(*p);
    const unsigned blocks = (unsigned)((p->num_envs + kThreads - 1) / kThreads);
    if (p->out_f64)
        fe_lazy_kernel<double, true><<<blocks, kThreads, 0, (cudaStream_t)stream>>>(
            *p, *s, *st, k, nullptr, obs_row0_dev, (double *)obs_posfeat_dev, nullptr, nullptr, nullptr, 0, nullptr);
    else
        fe_lazy_kernel<float, true><<<blocks, kThreads, 0, (cudaStream_t)stream>>>(
            *p, *s, *st, k, nullptr, obs_row0_dev, (float *)obs_posfeat_dev, nullptr, nullptr, nullptr, 0, nullptr);
    return (int)cudaGetLastError();
}

int fe_step_lazy(const FeParams *p, const FeSeries *s, const FeState *st, const float *actions_dev, int64_t *obs_row0_dev,
                 void *obs_posfeat_dev, void *rewards_dev, int32_t *dones_dev, FeStats *stats_dev, uint64_t step_counter,
                 void *stream) {
    int rc = check_common(p, s, st);
    if (rc) return rc;
    if (!actions_dev || !obs_row0_dev || !obs_posfeat_dev || !rewards_dev || !dones_dev || p->num_assets != 1) return FE_EINVAL;
    if (stats_dev && !p->evaluate && (!st->ep_return || !st->ep_len)) return FE_EINVAL;
    if ((rc = set_device(p->device))) return rc;
    const Consts k = make_consts(*p);
    const unsigned blocks = (unsigned)((p->num_envs + kThreads - 1) / kThreads);
    if (p->out_f64)
        fe_lazy_kernel<double, false><<<blocks, kThreads, 0, (cudaStream_t)stream>>>(
            *p, *s, *st, k, actions_dev, obs_row0_dev, (double *)obs_posfeat_dev, (double *)rewards_dev, dones_dev, stats_dev,
            step_counter, nullptr);
    else
        fe_lazy_kernel<float, false><<<blocks, kThreads, 0, (cudaStream_t)stream>>>(
            *p, *s, *st, k, actions_dev, obs_row0_dev, (float *)obs_posfeat_dev, (float *)rewards_dev, dones_dev, stats_dev,
            step_counter, nullptr);
    return (int)cudaGetLastError();
}

int fe_materialize(const FeParams *p, const FeSeries *s, const int64_t *obs_row0_dev, const void *obs_posfeat_dev,
                   void *obs_dev, void *stream) {
    if (!p || !s || !s->logret || !obs_row0_dev || !obs_posfeat_dev || !obs_dev) return FE_EINVAL;
    if (p->num_envs <= 0 || p->window <= 0 || p->num_assets != 1) return FE_EINVAL;
    int rc = set_device(p->device);
    if (rc) return rc;
    const unsigned blocks = (unsigned)((p->num_envs * 32 + kThreads - 1) / kThreads);
    if (p->out_f64)
        fe_materialize_kernel<double><<<blocks, kThreads, 0, (cudaStream_t)stream>>>(
            p->num_envs, p->window, (const double *)s->logret, obs_row0_dev, (const double *)obs_posfeat_dev, (double *)obs_dev);
    else
        fe_materialize_kernel<float><<<blocks, kThreads, 0, (cudaStream_t)stream>>>(
            p->num_envs, p->window, (const float *)s->logret, obs_row0_dev, (const float *)obs_posfeat_dev, (float *)obs_dev);
    return (int)cudaGetLastError();
}

// Host-buffer step, pipelined: the envs are cut into chunks (multiples of 1024 envs, so every chunk's slice of
// every array keeps its alignment); chunk c's action upload, kernel and result download run on side stream
// c % kHostStreams, so the upload of chunk c+1 and the download of chunk c-1 overlap the kernel of chunk c
// (two copy engines + SMs busy at once).  Redraws are keyed by global env id, so chunking changes no result.
int fe_step_host(const FeParams *p, const FeSeries *s, const FeState *st, const float *actions_host, float *actions_dev,
                 void *obs_dev, void *rewards_dev, int32_t *dones_dev, void *rewards_host, int32_t *dones_host,
                 FeStats *stats_dev, uint64_t step_counter, void *stream) {
    if (!p || !s || !st || !actions_host || !actions_dev || !rewards_host || !dones_host) return FE_EINVAL;
    int rc = check_common(p, s, st);
    if (rc) return rc;
    if (!obs_dev || !rewards_dev || !dones_dev) return FE_EINVAL;
    if (stats_dev && !p->evaluate && (!st->ep_return || !st->ep_len)) return FE_EINVAL;
    if ((rc = set_device(p->device))) return rc;
    cudaStream_t q = (cudaStream_t)stream;
    const int64_t n = p->num_envs;
    const int A = p->num_assets;
    const size_t osz = p->out_f64 ? 8 : 4;
    // Zero-copy mode: when all three host buffers are pinned (mapped into the device's address space), the step
    // kernel reads the actions from them and writes rewards / dones to them directly over PCIe — 12 bytes per env
    // spread over the whole kernel, no copy-engine hop, no head (upload) or tail (download) outside the kernel.
    // rewards_dev / dones_dev are written as well (device-side consumers); actions_dev is left untouched.
    // (Tried instead: copy-engine upload of the actions in chunks behind the already running kernel, each chunk
    // followed by a 4-byte copy bumping an arrival counter the envs poll.  Same speed on one GPU (0.317 ms per
    // 1 Mi-env step) and on eight (0.67 vs 0.68 ms: with 8 ranks the host side of PCIe, not the read latency, is the
    // limit), more machinery, and a kernel that waits for copies queued after it deadlocks under anything that
    // serialises launches (ncu, CUDA_LAUNCH_BLOCKING=1) — dropped.)
    static const int no_zc = env_override("FE_HOST_NO_ZEROCOPY");
    if (!no_zc) {
        cudaPointerAttributes aa, ar, ad;
        const bool ok = cudaPointerGetAttributes(&aa, actions_host) == cudaSuccess && aa.type == cudaMemoryTypeHost &&
                        aa.devicePointer && cudaPointerGetAttributes(&ar, rewards_host) == cudaSuccess &&
                        ar.type == cudaMemoryTypeHost && ar.devicePointer &&
                        cudaPointerGetAttributes(&ad, dones_host) == cudaSuccess && ad.type == cudaMemoryTypeHost &&
                        ad.devicePointer;
        (void)cudaGetLastError(); // an unregistered pointer leaves a sticky-free error on old drivers
        // Only for the persistent (and the portfolio) kernels: their bookkeeper warps run tiles ahead, which hides the
        // PCIe read latency, and they write rewards / dones 32 envs (128 bytes) per store.  A tile / direct block would
        // sit on its shared memory while it waits for its actions and write 16-byte PCIe packets (measured W = 60,
        // 1 Mi envs: 1.04 ms zero-copy, 0.77 ms with a copy-engine upload + zero-copy writes, vs 0.36 ms
        // device-resident): those take the chunked copy pipeline below.
        const StepKernel kk = choose_kernel(*p, p->out_f64 != 0).kern;
        if (ok && kk != K_TILE && kk != K_DIRECT) {
            const float *a = (const float *)aa.devicePointer;
            rc = p->out_f64 ? launch<double, false>(*p, *s, *st, a, obs_dev, rewards_dev, dones_dev, stats_dev, step_counter, q,
                                                    nullptr, ar.devicePointer, (int32_t *)ad.devicePointer)
                            : launch<float, false>(*p, *s, *st, a, obs_dev, rewards_dev, dones_dev, stats_dev, step_counter, q,
                                                   nullptr, ar.devicePointer, (int32_t *)ad.devicePointer);
            if (rc) return rc;
            return (int)cudaStreamSynchronize(q);
        }
    }
    static const int ov_chunks = env_override("FE_HOST_CHUNKS");
    int chunks = ov_chunks > 0 ? ov_chunks : 4;
    int64_t per = ((n + chunks - 1) / chunks + 1023) & ~(int64_t)1023;
    if (per < 16384) per = 16384; // below this a chunk's kernel is shorter than the launch + copy set-up it would hide
    chunks = (int)((n + per - 1) / per);
    HostPipe *hp = nullptr;
    if (chunks > 1) {
        hp = host_pipe(p->device);
        if (!hp) return (int)cudaGetLastError();
        cudaError_t e = cudaEventRecord(hp->start, q);
        if (e != cudaSuccess) return (int)e;
    }
    for (int c = 0; c < chunks; ++c) {
        const int64_t off = (int64_t)c * per, cnt = off + per <= n ? per : n - off;
        cudaStream_t cs = hp ? hp->s[c % kHostStreams] : q;
        cudaError_t e;
        if (hp && c < kHostStreams && (e = cudaStreamWaitEvent(cs, hp->start, 0)) != cudaSuccess) return (int)e;
        e = cudaMemcpyAsync(actions_dev + off * A, actions_host + off * A, (size_t)cnt * A * sizeof(float),
                            cudaMemcpyHostToDevice, cs);
        if (e != cudaSuccess) return (int)e;
        FeParams pc = *p;
        pc.num_envs = cnt;
        pc.env_id_base = p->env_id_base + off;
        FeState sc = *st;
        sc.seg += off; sc.ptr += off; sc.cash += off;
        sc.long_sh += off * A; sc.short_sh += off * A; sc.margin += off * A;
        if (sc.terminated) sc.terminated += off;
        if (sc.ep_return) sc.ep_return += off;
        if (sc.ep_len) sc.ep_len += off;
        const size_t obs_off = (size_t)off * p->window * 5 * A * osz;
        rc = fe_step(&pc, s, &sc, actions_dev + off * A, (char *)obs_dev + obs_off, (char *)rewards_dev + off * osz,
                     dones_dev + off, stats_dev, step_counter, cs);
        if (rc) return rc;
        e = cudaMemcpyAsync((char *)rewards_host + off * osz, (char *)rewards_dev + off * osz, (size_t)cnt * osz,
                            cudaMemcpyDeviceToHost, cs);
        if (e != cudaSuccess) return (int)e;
        e = cudaMemcpyAsync(dones_host + off, dones_dev + off, (size_t)cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, cs);
        if (e != cudaSuccess) return (int)e;
    }
    if (hp) { // the caller's stream continues only after every side stream has finished
        for (int k = 0; k < kHostStreams && k < chunks; ++k) {
            cudaError_t e = cudaEventRecord(hp->done[k], hp->s[k]);
            if (e != cudaSuccess) return (int)e;
            if ((e = cudaStreamWaitEvent(q, hp->done[k], 0)) != cudaSuccess) return (int)e;
        }
    }
    return (int)cudaStreamSynchronize(q);
}

int fe_reset_all(const FeParams *p, const FeSeries *s, const FeState *st, uint64_t step_counter, int32_t redraw,
                 void *stream) {
    int rc = check_common(p, s, st);
    if (rc) return rc;
    if ((rc = set_device(p->device))) return rc;
    fe_reset_all_kernel<<<(unsigned)((p->num_envs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*p, *s, *st, step_counter,
                                                                                                 redraw);
    return (int)cudaGetLastError();
}

void fe_philox(uint64_t seed, uint64_t env_id, uint64_t step, uint32_t kind, uint32_t out[4]) {
    philox4x32_10(seed, env_id, step, kind, out);
}

} // extern "C"
