/*
 * fe_oracle.c — CPU ORACLE (test infrastructure, NOT product code).  See fe_oracle.h.
 *
 * Every function cites the reference lines it restates; file =
 * /root/reference/finenvs/environments/time_series_env.py unless noted.  The dtype of each
 * expression follows SURVEY.md Appendix A: torch in-place ops compute in the promoted dtype and
 * round once to the destination; `f32_tensor * python_float` multiplies in f32; torch.round is
 * half-to-even.  Build with -ffp-contract=off (torch never fuses a*b+c).
 */
#include "fe_oracle.h"

#include <math.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ Philox4x32-10 ---------- */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void feo_philox(uint64_t seed, uint64_t env_id, uint64_t step, uint32_t kind, uint32_t out[4]) {
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t c[4] = {(uint32_t)env_id, (uint32_t)(env_id >> 32), (uint32_t)step,
                     ((uint32_t)(step >> 32) & 0x7FFFFFFFu) | (kind << 31)};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof(c));
}

/* Replaces torch.randint(0, num_days) at :253 / :511 with a counter-based draw so that CPU oracle,
 * GPU kernel and (via monkeypatch) the reference consume the same numbers.  Multiply-high maps a
 * u32 onto [0, n). */
void feo_draw(const FeoParams *p, const FeoSeries *s, int64_t env_id, uint64_t step, uint32_t kind,
              int32_t *seg_out, int32_t *off_out) {
    uint32_t r[4];
    feo_philox(p->seed, (uint64_t)env_id, step, kind, r);
    const int32_t seg = (int32_t)(((uint64_t)r[0] * (uint64_t)(uint32_t)p->num_segments) >> 32);
    int32_t off = 0;
    if (p->random_offset) {
        /* valid start pointers are 0 .. seg_len - W - 1 (one bar must remain to step onto) */
        const int32_t span = s->seg_len[seg] - p->window;
        off = span > 0 ? (int32_t)(((uint64_t)r[1] * (uint64_t)(uint32_t)span) >> 32) : 0;
    }
    *seg_out = seg;
    *off_out = off;
}

/* :477-496 — an env is done when the NEXT row index (ptr+W) is >= padded length L or that row's
 * log-return col 0 is NaN (NaN padding of a short day, or a genuine NaN from a bad price).  In the
 * flat layout both collapse to "ptr + W >= effective length". */
int32_t feo_effective_len(const double *logret, int64_t seg_start, int32_t raw_len, int32_t window,
                          int32_t num_assets) {
    for (int32_t k = window + 1; k < raw_len; ++k) {
        if (isnan(logret[((size_t)(seg_start + k) * num_assets) * 4])) return k;
    }
    return raw_len;
}

/* :179-194 — cols 1..3 = 100*log(H|L|C / O); col 0 = 100*log(O_t / C_{t-1}), first row uses O_0. */
void feo_log_returns(const double *prices, int64_t num_rows, int32_t num_assets, double *logret,
                     float *logret32) {
    const int64_t A = num_assets;
    for (int64_t t = 0; t < num_rows; ++t) {
        for (int64_t a = 0; a < A; ++a) {
            const double *px = prices + (t * A + a) * 4;
            const double prev_close = t == 0 ? px[0] : prices[((t - 1) * A + a) * 4 + 3];
            double lr[4];
            lr[0] = 100.0 * log(px[0] / prev_close);
            for (int c = 1; c < 4; ++c) lr[c] = 100.0 * log(px[c] / px[0]);
            for (int c = 0; c < 4; ++c) {
                if (logret) logret[(t * A + a) * 4 + c] = lr[c];
                if (logret32) logret32[(t * A + a) * 4 + c] = (float)lr[c];
            }
        }
    }
}

/* torch.relu keeps NaN (clamp_min), unlike fmax */
static inline float relu32(float x) { return x < 0.0f ? 0.0f : x; }
static inline double relu64(double x) { return x < 0.0 ? 0.0 : x; }

/* :437-445 window gather + :428-434 position feature, one env */
static void write_obs(const FeoParams *p, const FeoSeries *s, void *obs, int64_t i, int64_t row0,
                      double posfeat) {
    const int W = p->window;
    if (p->out_f64) {
        double *o = (double *)obs + (size_t)i * W * 5;
        const double *src = s->logret + (size_t)row0 * 4;
        for (int j = 0; j < W; ++j) {
            o[5 * j + 0] = src[4 * j + 0];
            o[5 * j + 1] = src[4 * j + 1];
            o[5 * j + 2] = src[4 * j + 2];
            o[5 * j + 3] = src[4 * j + 3];
            o[5 * j + 4] = posfeat;
        }
    } else {
        float *o = (float *)obs + (size_t)i * W * 5;
        const float *src = s->logret32 + (size_t)row0 * 4;
        const float pf = (float)posfeat;
        for (int j = 0; j < W; ++j) {
            o[5 * j + 0] = src[4 * j + 0];
            o[5 * j + 1] = src[4 * j + 1];
            o[5 * j + 2] = src[4 * j + 2];
            o[5 * j + 3] = src[4 * j + 3];
            o[5 * j + 4] = pf;
        }
    }
}

/* reference reset() :423-435 — it never touches state, it only materialises the observation.
 * The close used is the one at row ptr+W-1 (:326 env_spots[:, -1]); for a just-reset env the
 * reference holds a stale close but shares are 0 there, so the product is 0 either way. */
void feo_observe(const FeoParams *p, const FeoSeries *s, const FeoState *st, void *obs) {
    const int W = p->window;
#pragma omp parallel for schedule(static) if (p->num_envs >= 4096)
    for (int64_t i = 0; i < p->num_envs; ++i) {
        const int64_t row0 = s->seg_start[st->seg[i]] + st->ptr[i];
        const double C = s->prices[(size_t)(row0 + W - 1) * 4 + 3];
        const float net = st->long_sh[i] - st->short_sh[i];                   /* f32 :428-429 */
        const double posfeat = ((double)net * C) / p->starting_balance;       /* f64 :430-431 */
        write_obs(p, s, obs, i, row0, posfeat);
    }
}

int64_t feo_step(const FeoParams *p, const FeoSeries *s, const FeoState *st, const float *actions,
                 void *obs, void *rewards, int32_t *dones, uint64_t step_counter,
                 int32_t *all_terminated) {
    const int W = p->window;
    const float ms = (float)p->max_shares;
    const float scale = (float)((double)p->max_shares + 0.5); /* :299 python float -> f32 scalar */
    const double c = p->commission;
    const float cf = (float)p->commission;                    /* f32 tensor * python float */
    const float imrf = (float)p->imr;
    const double imr = p->imr;
    const double mmr1 = 1.0 + p->mmr;                         /* :462 python-side (1 + mmr) */
    const double SB = p->starting_balance;
    const float SBf = (float)p->starting_balance;             /* :499 store into f32 cash */
    int64_t n_done = 0;
    int64_t n_not_terminated = 0;

#pragma omp parallel for schedule(static) reduction(+ : n_done, n_not_terminated) if (p->num_envs >= 4096)
    for (int64_t i = 0; i < p->num_envs; ++i) {
        /* :298-302 action -> integer share delta */
        float d = rintf(actions[i] * scale);
        d = d < -ms ? -ms : (d > ms ? ms : d);
        /* :281-282 advance time */
        int32_t seg = st->seg[i];
        int32_t ptr = st->ptr[i] + 1;
        const int64_t row0 = s->seg_start[seg] + ptr;
        /* :323-342 current bar = last row of the window */
        const double *px = s->prices + (size_t)(row0 + W - 1) * 4;
        const double O = px[0], H = px[1], L = px[2], C = px[3];
        float cash = st->cash[i];
        float lng = st->long_sh[i];
        float sht = st->short_sh[i];
        double margin = st->margin[i];
        float comm = 0.0f;                                     /* :305 */
        /* :344-351 */
        float pos = d < 0.0f ? 0.0f : d;
        float neg = d > 0.0f ? 0.0f : d;
        /* :353-361 sell longs */
        {
            const float nl = relu32(lng + neg);
            const float sold = lng - nl;
            neg = neg + sold;
            comm = comm + sold * cf;                           /* :363-365 */
            cash = (float)((double)cash + (double)sold * (O - c));
            lng = nl;
        }
        /* :367-383 cover shorts, re-mark margin */
        {
            const float ns = relu32(sht - pos);
            const float bought = sht - ns;
            pos = pos - bought;
            comm = comm + bought * cf;
            cash = (float)((double)cash - (double)bought * (O + c));
            sht = ns;
            const double nm = (double)(imrf * sht) * O;        /* :376-380 (imr*short) in f32 */
            cash = (float)((double)cash - (nm - margin));
            margin = nm;
        }
        /* :385-392 all-or-nothing long entry */
        if (((double)cash - (double)pos * (O + c)) < 0.0) pos = 0.0f;
        /* :394-399 */
        comm = comm + pos * cf;
        cash = (float)((double)cash - (double)pos * (O + c));
        lng = lng + pos;
        /* :401-410 all-or-nothing short entry */
        {
            float q = -neg;
            float sc = q * cf;
            double req = imr * ((double)q * O);
            if ((((double)cash - req) - (double)sc) < 0.0) {
                neg = 0.0f;
                q = -neg;
                sc = q * cf;
                req = imr * ((double)q * O);
            }
            /* :412-421 */
            comm = comm + q * cf;
            cash = (float)((double)cash - (req + (double)sc));
            margin = margin + req;
            sht = sht + q;
        }
        /* :321 -> reset(): the observation is built NOW (before rewards / dones / auto-reset) */
        if (obs) {
            const float net = lng - sht;
            write_obs(p, s, obs, i, row0, ((double)net * C) / SB);
        }
        /* :447-457 rewards */
        int done = cash < 0.0f;                                /* :448 */
        double rew;
        {
            /* :459-468 maintenance margin at High */
            const double mc1 = relu64(((double)sht * H) * mmr1 - margin);
            cash = (float)((double)cash - mc1);
            margin = margin + mc1;
            done |= cash < 0.0f;
            /* :470-475 margin release at Low */
            const double rel = relu64(margin - ((double)sht * L) * imr);
            margin = margin - rel;
            cash = (float)((double)cash + rel);
            /* :451 maintenance margin at Close */
            const double mc2 = relu64(((double)sht * C) * mmr1 - margin);
            cash = (float)((double)cash - mc2);
            margin = margin + mc2;
            done |= cash < 0.0f;
            rew = (-mc1) + (-mc2);
            if (done) { lng = 0.0f; sht = 0.0f; }              /* :452-453 */
            rew = rew + (double)(lng - sht) * (C - O);         /* :454-455 */
            rew = rew - (double)comm;                          /* :456 */
        }
        /* :477-496 time limit / NaN padding */
        done |= (ptr + W >= s->seg_len[seg]);
        /* :288-289 closing commission on whatever is still held */
        rew = rew - (double)(((done ? 1.0f : 0.0f) * (sht + lng)) * cf);
        /* :498-521 auto-reset */
        if (done) {
            cash = SBf; margin = 0.0; lng = 0.0f; sht = 0.0f; ptr = 0;
            const int64_t gid = p->env_id_base + i;
            const int redraw = p->reset_mode == FEO_RESET_ALL ||
                               (p->reset_mode == FEO_RESET_LAST && gid == p->total_envs - 1);
            if (redraw) {
                int32_t nseg, off;
                feo_draw(p, s, gid, step_counter, 0u, &nseg, &off);
                seg = nseg;
                ptr = off;
            }
            ++n_done;
        }
        /* :523-536 evaluate bookkeeping */
        if (p->evaluate) {
            if (st->terminated[i]) rew = 0.0;                  /* :527-528 */
            if (done) st->terminated[i] = 1;                   /* :529 */
            st->ep_return[i] = (float)((double)st->ep_return[i] + rew); /* :530 f32 += f64 */
            if (!st->terminated[i]) ++n_not_terminated;
        }
        st->seg[i] = seg; st->ptr[i] = ptr; st->cash[i] = cash;
        st->long_sh[i] = lng; st->short_sh[i] = sht; st->margin[i] = margin;
        if (p->out_f64) ((double *)rewards)[i] = rew; else ((float *)rewards)[i] = (float)rew;
        dones[i] = done;                                       /* :296 dones.int() */
    }
    if (all_terminated) *all_terminated = p->evaluate && n_not_terminated == 0;
    return n_done;
}

/* ------------------------------------------------------------------ multi-asset (A > 1) -----
 * EXTENSION with no reference implementation (the reference is single-asset, :223); SURVEY App. D.
 * Definition: shared f32 cash, per-asset long/short/margin; the reference's phases (one per cash update
 * of the reference: sell longs, cover shorts, re-mark margin, long entries, short entries, margin call at
 * High, release at Low, margin call at Close) run in the reference's order, and in each phase cash moves
 * ONCE (f64 arithmetic, one rounding to f32): where the per-asset deltas do not depend on cash, by their
 * butterfly sum (bankruptcy is tested on the result); in the two entry phases, by a greedy walk over the
 * assets in index order with a running f64 cash — an entry is legal if the cash left by the assets before
 * it pays for it.  A = 1 reduces to feo_step() exactly (tested on every golden trace).  Per-env sums over assets (reward, share count) use the xor-butterfly order of a warp
 * reduction so that the CUDA kernel can reproduce them bit for bit. */
#define FEO_MAX_ASSETS 32

static double butterfly_sum64(const double *v, int A) {
    double w[32];
    for (int i = 0; i < 32; ++i) w[i] = i < A ? v[i] : 0.0;
    for (int off = 16; off > 0; off >>= 1) {
        double n[32];
        for (int i = 0; i < 32; ++i) n[i] = w[i] + w[i ^ off];
        memcpy(w, n, sizeof(w));
    }
    return w[0];
}

static float butterfly_sum32(const float *v, int A) {
    float w[32];
    for (int i = 0; i < 32; ++i) w[i] = i < A ? v[i] : 0.0f;
    for (int off = 16; off > 0; off >>= 1) {
        float n[32];
        for (int i = 0; i < 32; ++i) n[i] = w[i] + w[i ^ off];
        memcpy(w, n, sizeof(w));
    }
    return w[0];
}

/* obs (N, W, A, 5): row j, asset a = [logret[row0+j, a, 0..3], posfeat_a] */
static void write_obs_multi(const FeoParams *p, const FeoSeries *s, void *obs, int64_t i, int64_t row0,
                            const double *posfeat) {
    const int W = p->window, A = p->num_assets;
    for (int j = 0; j < W; ++j)
        for (int a = 0; a < A; ++a) {
            const size_t src = ((size_t)(row0 + j) * A + a) * 4;
            const size_t dst = (((size_t)i * W + j) * A + a) * 5;
            if (p->out_f64) {
                double *o = (double *)obs + dst;
                for (int c = 0; c < 4; ++c) o[c] = s->logret[src + c];
                o[4] = posfeat[a];
            } else {
                float *o = (float *)obs + dst;
                for (int c = 0; c < 4; ++c) o[c] = s->logret32[src + c];
                o[4] = (float)posfeat[a];
            }
        }
}

void feo_observe_multi(const FeoParams *p, const FeoSeries *s, const FeoState *st, void *obs) {
    const int W = p->window, A = p->num_assets;
#pragma omp parallel for schedule(static) if (p->num_envs >= 1024)
    for (int64_t i = 0; i < p->num_envs; ++i) {
        const int64_t row0 = s->seg_start[st->seg[i]] + st->ptr[i];
        double pf[FEO_MAX_ASSETS];
        for (int a = 0; a < A; ++a) {
            const double C = s->prices[((size_t)(row0 + W - 1) * A + a) * 4 + 3];
            const float net = st->long_sh[i * A + a] - st->short_sh[i * A + a];
            pf[a] = ((double)net * C) / p->starting_balance;
        }
        write_obs_multi(p, s, obs, i, row0, pf);
    }
}

int64_t feo_step_multi(const FeoParams *p, const FeoSeries *s, const FeoState *st, const float *actions,
                       void *obs, void *rewards, int32_t *dones, uint64_t step_counter,
                       int32_t *all_terminated) {
    const int W = p->window, A = p->num_assets;
    const float ms = (float)p->max_shares;
    const float scale = (float)((double)p->max_shares + 0.5);
    const double c = p->commission;
    const float cf = (float)p->commission;
    const float imrf = (float)p->imr;
    const double imr = p->imr;
    const double mmr1 = 1.0 + p->mmr;
    const double SB = p->starting_balance;
    const float SBf = (float)p->starting_balance;
    int64_t n_done = 0, n_not_terminated = 0;

#pragma omp parallel for schedule(static) reduction(+ : n_done, n_not_terminated) if (p->num_envs >= 1024)
    for (int64_t i = 0; i < p->num_envs; ++i) {
        float pos[FEO_MAX_ASSETS], neg[FEO_MAX_ASSETS], lng[FEO_MAX_ASSETS], sht[FEO_MAX_ASSETS], comm[FEO_MAX_ASSETS];
        double margin[FEO_MAX_ASSETS], O[FEO_MAX_ASSETS], H[FEO_MAX_ASSETS], L[FEO_MAX_ASSETS], C[FEO_MAX_ASSETS];
        double mc1[FEO_MAX_ASSETS], mc2[FEO_MAX_ASSETS], r[FEO_MAX_ASSETS], pf[FEO_MAX_ASSETS];
        int32_t seg = st->seg[i];
        int32_t ptr = st->ptr[i] + 1;
        const int64_t row0 = s->seg_start[seg] + ptr;
        float cash = st->cash[i];
        for (int a = 0; a < A; ++a) {
            float d = rintf(actions[i * A + a] * scale);                 /* :298-302 */
            d = d < -ms ? -ms : (d > ms ? ms : d);
            pos[a] = d < 0.0f ? 0.0f : d;                                /* :344-351 */
            neg[a] = d > 0.0f ? 0.0f : d;
            const double *px = s->prices + ((size_t)(row0 + W - 1) * A + a) * 4;
            O[a] = px[0]; H[a] = px[1]; L[a] = px[2]; C[a] = px[3];
            lng[a] = st->long_sh[i * A + a]; sht[a] = st->short_sh[i * A + a]; margin[a] = st->margin[i * A + a];
            comm[a] = 0.0f;
        }
        double dl[FEO_MAX_ASSETS];                                       /* per-asset cash deltas of a phase */
        double c64;
        for (int a = 0; a < A; ++a) {                                    /* :353-361 */
            const float nl = relu32(lng[a] + neg[a]);
            const float sold = lng[a] - nl;
            neg[a] = neg[a] + sold;
            comm[a] = comm[a] + sold * cf;
            dl[a] = (double)sold * (O[a] - c);
            lng[a] = nl;
        }
        cash = (float)((double)cash + butterfly_sum64(dl, A));
        for (int a = 0; a < A; ++a) {                                    /* :367-374 */
            const float ns = relu32(sht[a] - pos[a]);
            const float bought = sht[a] - ns;
            pos[a] = pos[a] - bought;
            comm[a] = comm[a] + bought * cf;
            dl[a] = (double)bought * (O[a] + c);
            sht[a] = ns;
        }
        cash = (float)((double)cash - butterfly_sum64(dl, A));
        for (int a = 0; a < A; ++a) {                                    /* :375-383 */
            const double nm = (double)(imrf * sht[a]) * O[a];
            dl[a] = nm - margin[a];
            margin[a] = nm;
        }
        cash = (float)((double)cash - butterfly_sum64(dl, A));
        c64 = (double)cash;
        for (int a = 0; a < A; ++a) {                                    /* :385-399 */
            const double cost = (double)pos[a] * (O[a] + c);
            if ((c64 - cost) < 0.0) pos[a] = 0.0f;
            else c64 = c64 - cost;
            comm[a] = comm[a] + pos[a] * cf;
            lng[a] = lng[a] + pos[a];
        }
        cash = (float)c64;
        c64 = (double)cash;
        for (int a = 0; a < A; ++a) {                                    /* :401-421 */
            float q = -neg[a];
            float sc = q * cf;
            double req = imr * ((double)q * O[a]);
            if (((c64 - req) - (double)sc) < 0.0) {
                q = -0.0f;
                sc = q * cf;
                req = imr * ((double)q * O[a]);
            } else {
                c64 = c64 - (req + (double)sc);
            }
            comm[a] = comm[a] + q * cf;
            margin[a] = margin[a] + req;
            sht[a] = sht[a] + q;
        }
        cash = (float)c64;
        for (int a = 0; a < A; ++a) pf[a] = ((double)(lng[a] - sht[a]) * C[a]) / SB;   /* :428-431 */
        if (obs) write_obs_multi(p, s, obs, i, row0, pf);
        int done = cash < 0.0f;                                          /* :448 */
        for (int a = 0; a < A; ++a) {                                    /* :459-468 at High */
            mc1[a] = relu64(((double)sht[a] * H[a]) * mmr1 - margin[a]);
            margin[a] = margin[a] + mc1[a];
        }
        cash = (float)((double)cash - butterfly_sum64(mc1, A));
        done |= cash < 0.0f;
        for (int a = 0; a < A; ++a) {                                    /* :470-475 at Low */
            dl[a] = relu64(margin[a] - ((double)sht[a] * L[a]) * imr);
            margin[a] = margin[a] - dl[a];
        }
        cash = (float)((double)cash + butterfly_sum64(dl, A));
        for (int a = 0; a < A; ++a) {                                    /* :451 at Close */
            mc2[a] = relu64(((double)sht[a] * C[a]) * mmr1 - margin[a]);
            margin[a] = margin[a] + mc2[a];
        }
        cash = (float)((double)cash - butterfly_sum64(mc2, A));
        done |= cash < 0.0f;
        float held[FEO_MAX_ASSETS];
        for (int a = 0; a < A; ++a) {
            double ra = (-mc1[a]) + (-mc2[a]);
            if (done) { lng[a] = 0.0f; sht[a] = 0.0f; }                  /* :452-453 */
            ra = ra + (double)(lng[a] - sht[a]) * (C[a] - O[a]);         /* :454-455 */
            r[a] = ra - (double)comm[a];                                 /* :456 */
            held[a] = sht[a] + lng[a];                                   /* :288 */
        }
        double rew = butterfly_sum64(r, A);
        const float nsh = butterfly_sum32(held, A);
        done |= (ptr + W >= s->seg_len[seg]);                            /* :477-496 */
        rew = rew - (double)(((done ? 1.0f : 0.0f) * nsh) * cf);         /* :288-289 */
        if (done) {                                                      /* :498-521 */
            cash = SBf; ptr = 0;
            for (int a = 0; a < A; ++a) { margin[a] = 0.0; lng[a] = 0.0f; sht[a] = 0.0f; }
            const int64_t gid = p->env_id_base + i;
            if (p->reset_mode == FEO_RESET_ALL || (p->reset_mode == FEO_RESET_LAST && gid == p->total_envs - 1)) {
                int32_t nseg, off;
                feo_draw(p, s, gid, step_counter, 0u, &nseg, &off);
                seg = nseg; ptr = off;
            }
            ++n_done;
        }
        if (p->evaluate) {                                               /* :523-536 */
            if (st->terminated[i]) rew = 0.0;
            if (done) st->terminated[i] = 1;
            st->ep_return[i] = (float)((double)st->ep_return[i] + rew);
            if (!st->terminated[i]) ++n_not_terminated;
        }
        st->seg[i] = seg; st->ptr[i] = ptr; st->cash[i] = cash;
        for (int a = 0; a < A; ++a) {
            st->long_sh[i * A + a] = lng[a]; st->short_sh[i * A + a] = sht[a]; st->margin[i * A + a] = margin[a];
        }
        if (p->out_f64) ((double *)rewards)[i] = rew; else ((float *)rewards)[i] = (float)rew;
        dones[i] = done;
    }
    if (all_terminated) *all_terminated = p->evaluate && n_not_terminated == 0;
    return n_done;
}

/* Extension: start a fresh episode everywhere; optional redraw of (segment, offset).
 * Mirrors the state :245-269 allocates. */
void feo_reset_all(const FeoParams *p, const FeoSeries *s, const FeoState *st, uint64_t step_counter,
                   int32_t redraw) {
    const float SBf = (float)p->starting_balance;
#pragma omp parallel for schedule(static) if (p->num_envs >= 4096)
    for (int64_t i = 0; i < p->num_envs; ++i) {
        st->cash[i] = SBf;
        for (int a = 0; a < p->num_assets; ++a) {
            st->margin[i * p->num_assets + a] = 0.0;
            st->long_sh[i * p->num_assets + a] = 0.0f;
            st->short_sh[i * p->num_assets + a] = 0.0f;
        }
        st->ptr[i] = 0;
        if (redraw) {
            int32_t seg, off;
            feo_draw(p, s, p->env_id_base + i, step_counter, 1u, &seg, &off);
            st->seg[i] = seg; st->ptr[i] = off;
        }
        if (p->evaluate) { st->terminated[i] = 0; st->ep_return[i] = 0.0f; }
    }
}

int feo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void feo_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
