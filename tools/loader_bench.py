#!/usr/bin/env python
"""Loader & staging: finenvs_b200.data.loader against the reference's process_data().

    python tools/loader_bench.py [--days 256] [--bars 390] [--window 60]          # CPU part (runs anywhere)
    python tools/loader_bench.py --stage-rows 10000000                            # + GPU staging of a flat series

CPU part: a reference-format CSV (no header; Date,Time,Open,High,Low,Close,Volume; one Date per segment, bar j at 09:30 + j min)
is written to a scratch directory and loaded by (a) loader.read_market_csv — one vectorised pass: day-change detection ->
flat series + segment table — and, where the reference checkout exists (this container only), (b) the reference's
own constructor path (time_series_env.py:80-216: pandas between_time + one boolean scan of the whole frame per day),
imported read-only through oracle/ref_harness.  GPU part: loader.stage_series of synthetic rows (pinned upload,
fe_log_returns, fe_effective_len).  Prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def write_csv(path, days, bars, rng):
    from parity_utils import gbm_ohlc

    px = np.round(gbm_ohlc(rng, days * bars, 0.0005), 4)
    d0 = np.datetime64("2020-01-01")
    with open(path, "w") as f:
        for d in range(days):
            date = str(d0 + d).replace("-", "/")
            date = f"{date[5:7]}/{date[8:10]}/{date[0:4]}"
            for j in range(bars):
                m = 9 * 60 + 30 + j
                o, h, l, c = px[d * bars + j]
                f.write(f"{date},{m // 60:02d}:{m % 60:02d},{o:.4f},{h:.4f},{l:.4f},{c:.4f},100\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--days", type=int, default=256)
    ap.add_argument("--bars", type=int, default=390)
    ap.add_argument("--window", type=int, default=60)
    ap.add_argument("--stage-rows", type=int, default=0)
    ap.add_argument("--no-reference", action="store_true", help="skip the reference constructor (O(days x rows): hours for millions of rows)")
    args = ap.parse_args()
    from finenvs_b200.data import loader

    out = {"csv": f"{args.days} days x {args.bars} one-minute bars", "window": args.window}
    tmp = tempfile.mkdtemp(prefix="fe_loader_")
    try:
        sub = os.path.join(tmp, "data", "SYN")
        os.makedirs(sub)
        path = os.path.join(sub, "dummy.csv")
        write_csv(path, args.days, args.bars, np.random.default_rng(0))
        import pandas  # noqa: F401  (not part of the timed region, as for the reference below)

        t0 = time.perf_counter()
        host = loader.read_market_csv(path, args.window)            # native reader (csrc/fe_csv.cu), all host threads
        out["loader_read_market_csv_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        loader.read_market_csv(path, args.window, reader="pandas")  # the same loader on pandas.read_csv
        out["loader_read_market_csv_pandas_s"] = time.perf_counter() - t0
        out["host_threads"] = os.cpu_count()
        out["segments"] = int(len(host.seg_start))
        out["rows"] = int(host.prices.shape[0])
        try:
            from oracle import ref_harness

            if ref_harness.available() and not args.no_reference:
                mod = ref_harness.ref_module()
                t0 = time.perf_counter()
                env = mod.TimeSeriesEnv(sub, "dummy", num_intervals=args.window, device_id=-1)
                out["reference_constructor_s"] = time.perf_counter() - t0
                out["reference_segments"] = int(env.price_environments.shape[0])
        except Exception as e:  # the checkout is optional
            out["reference_error"] = repr(e)[:200]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)

    if args.stage_rows:
        import torch
        from parity_utils import gbm_ohlc

        T = args.stage_rows
        prices = np.round(gbm_ohlc(np.random.default_rng(1), T, 0.0005), 4)
        seg_start, seg_len = loader.regular_segments(T, 390, args.window)
        loader.stage_series(prices[:100000], *loader.regular_segments(100000, 390, args.window), args.window, "cuda:0", torch.float32)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        series = loader.stage_series(prices, seg_start, seg_len, args.window, "cuda:0", torch.float32)
        torch.cuda.synchronize()
        out["stage_series_s"] = time.perf_counter() - t0
        out["staged_rows"] = T
        out["staged_bytes"] = int(series.prices.numel() * 8 + series.logret.numel() * 4)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
