/*
 * fe_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A scalar, per-env C restatement of the reference's vectorised trading-env step
 * (hmomin/FinEnvs finenvs/environments/time_series_env.py:277-536) working on the
 * flat-series + segment-table layout the B200 build uses.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this.  The product path (finenvs_b200/) never does.
 *
 * Parity pin: checked bit-for-bit against the imported reference (tests/test_oracle_vs_reference.py,
 * run where /root/reference exists) and against the committed golden traces that
 * the reference itself produced (tests/golden/ npz files, made by tests/golden/make_golden.py).
 */
#ifndef FE_ORACLE_H
#define FE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reset_mode */
#define FEO_RESET_KEEP 0 /* done envs restart on the same segment (reference evaluate=True, :504) */
#define FEO_RESET_LAST 1 /* only the globally-last env redraws its segment (reference training, :504-513) */
#define FEO_RESET_ALL 2  /* extension: every done env redraws its segment */

typedef struct FeoParams {
    int64_t num_envs;      /* envs held by this call (a shard) */
    int64_t env_id_base;   /* global id of env 0 of this shard */
    int64_t total_envs;    /* global env count (the "last env" is total_envs-1) */
    int64_t num_rows;      /* T rows in the flat series */
    int32_t window;        /* W = num_intervals */
    int32_t num_segments;  /* D */
    int32_t num_assets;    /* A (1 = the reference env) */
    int32_t max_shares;
    double starting_balance;
    double commission;     /* per_share_commission */
    double imr;            /* initial_margin_requirement */
    double mmr;            /* maintenance_margin_requirement */
    uint64_t seed;
    int32_t reset_mode;
    int32_t random_offset; /* extension: redraw also draws a start offset inside the segment */
    int32_t evaluate;      /* reference evaluate=True bookkeeping (:523-536) */
    int32_t out_f64;       /* 1: obs/rewards written as double (reference dtype); 0: float */
} FeoParams;

typedef struct FeoSeries {
    const double *prices;    /* (T, A, 4) OHLC f64 */
    const double *logret;    /* (T, A, 4) 100*log-returns f64 (obs source when out_f64) */
    const float *logret32;   /* (T, A, 4) same rounded to f32 (obs source when !out_f64) */
    const int64_t *seg_start; /* (D,) first row of each segment (W history rows included) */
    const int32_t *seg_len;   /* (D,) EFFECTIVE rows in the segment (see feo_effective_len) */
} FeoSeries;

typedef struct FeoState {
    int32_t *seg;      /* (N,)  env -> segment (reference env_indices :246) */
    int32_t *ptr;      /* (N,)  time pointer (reference env_pointers :258; env_spots[i,j] == ptr[i]+j) */
    float *cash;       /* (N,)   :264 */
    float *long_sh;    /* (N,A)  :267 */
    float *short_sh;   /* (N,A)  :268 */
    double *margin;    /* (N,A)  :269 (f64 after the first step, :383) */
    uint8_t *terminated; /* (N,) evaluate only :272 */
    float *ep_return;    /* (N,) evaluate only :275 */
} FeoState;

/* Philox4x32-10, key=(seed lo,hi), counter=(env lo, env hi, step lo, step hi|kind<<31). */
void feo_philox(uint64_t seed, uint64_t env_id, uint64_t step, uint32_t kind, uint32_t out[4]);

/* segment / offset draw used by every redraw (kind 0 = step-time reset, 1 = reset_all/constructor) */
void feo_draw(const FeoParams *p, const FeoSeries *s, int64_t env_id, uint64_t step, uint32_t kind,
              int32_t *seg_out, int32_t *off_out);

/* rows of segment d before the reference's NaN probe (:486-496) would fire:
 * min(raw_len, first k >= W+1 with isnan(logret[start+k, 0, 0])). */
int32_t feo_effective_len(const double *logret, int64_t seg_start, int32_t raw_len, int32_t window,
                          int32_t num_assets);

/* 100*log-return table from OHLC (reference :179-194). */
void feo_log_returns(const double *prices, int64_t num_rows, int32_t num_assets, double *logret,
                     float *logret32);

/* reference reset() (:423-445): writes obs only; obs is (N, W, 5A) float or double */
void feo_observe(const FeoParams *p, const FeoSeries *s, const FeoState *st, void *obs);

/* reference step() (:277-296). actions (N,A) f32; obs (N,W,5A); rewards (N,); dones (N,) i32.
 * step_counter = number of this step (1 for the first step after construction).
 * obs may be NULL (skip the window write).  Returns number of envs that finished.
 * all_terminated (may be NULL): evaluate mode, set to 1 when every env of this shard has terminated. */
int64_t feo_step(const FeoParams *p, const FeoSeries *s, const FeoState *st, const float *actions,
                 void *obs, void *rewards, int32_t *dones, uint64_t step_counter, int32_t *all_terminated);

/* EXTENSION without a reference implementation: A > 1 assets sharing one cash account (A <= 32);
 * series are time-major (T, A, 4), obs is (N, W, A, 5).  At A = 1 identical to feo_step/feo_observe. */
void feo_observe_multi(const FeoParams *p, const FeoSeries *s, const FeoState *st, void *obs);
int64_t feo_step_multi(const FeoParams *p, const FeoSeries *s, const FeoState *st, const float *actions,
                       void *obs, void *rewards, int32_t *dones, uint64_t step_counter, int32_t *all_terminated);

/* extension: fresh episode for every env (mirrors isaac_gym_env.py:55-58 reset_all). */
void feo_reset_all(const FeoParams *p, const FeoSeries *s, const FeoState *st, uint64_t step_counter,
                   int32_t redraw);

int feo_num_threads(void);
void feo_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
