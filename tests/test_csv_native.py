"""The library's CSV reader (csrc/fe_csv.cu: fe_csv_open / fe_csv_read / fe_csv_close) against pandas.read_csv — the
reader the reference uses (time_series_env.py:80-88).  Prices must be BIT-identical (the staged series feeds the
bit-exact bookkeeping), times and day grouping equal, and anything outside the native format must fall back.  CPU only."""
import os

import numpy as np
import pytest

from finenvs_b200.data import loader


def _rand_number(rng):
    kind = int(rng.integers(0, 10))
    ip = str(int(rng.integers(0, 10 ** int(rng.integers(1, 7)))))
    if kind == 0:
        return ip                                                   # integer-looking price
    nd = int(rng.integers(1, 10)) if kind < 6 else int(rng.integers(10, 24))
    s = ip + "." + "".join(str(int(d)) for d in rng.integers(0, 10, nd))
    if kind == 7:                                                   # more than 17 significant digits before the point
        s = str(int(rng.integers(1, 10 ** 18))) + str(int(rng.integers(0, 10 ** 6))) + "." + s[-3:]
    if kind == 8:                                                   # scientific notation, sign
        s = ("-" if rng.integers(0, 2) else "+") + s + "eE"[int(rng.integers(0, 2))] + str(int(rng.integers(-12, 13)))
    if kind == 9:
        s = "000" + s                                               # leading zeros count as digits in pandas' converter
    return s


def _write(path, n, rng, newline="\n", seconds=False, trailing_newline=True, blank_every=0):
    lines = []
    for i in range(n):
        date = f"{1 + (i // 400) % 12:02d}/{1 + (i // 40) % 28:02d}/{1998 + i // 100000}"
        time = f"{9 + (i % 400) // 60:02d}:{i % 60:02d}" + (f":{(7 * i) % 60:02d}" if seconds else "")
        lines.append(",".join([date, time] + [_rand_number(rng) for _ in range(4)] + [str(int(rng.integers(0, 10 ** 7)))]))
        if blank_every and i % blank_every == 0:
            lines.append("")
    with open(path, "w", newline="") as f:
        f.write(newline.join(lines) + (newline if trailing_newline else ""))
    return lines


def _assert_same(nat, pan):
    import pandas as pd

    assert nat is not None
    assert nat[2].shape == pan[2].shape
    assert np.array_equal(nat[2].view(np.int64), pan[2].view(np.int64))          # prices: same bits
    assert np.array_equal(nat[1], pan[1])                                        # seconds of day
    assert np.array_equal(pd.factorize(nat[0])[0], pd.factorize(pan[0])[0])      # same rows share a day, same order


@pytest.mark.parametrize("seed,kw", [(1, {}), (2, dict(newline="\r\n")), (3, dict(seconds=True, trailing_newline=False)),
                                     (4, dict(blank_every=7))])
def test_native_reader_equals_pandas_bit_for_bit(tmp_path, seed, kw):
    p = str(tmp_path / "m.csv")
    _write(p, 40000, np.random.default_rng(seed), **kw)
    _assert_same(loader.read_csv_native(p), loader.read_csv_pandas(p))


def test_thread_count_does_not_change_the_result(tmp_path):
    p = str(tmp_path / "m.csv")
    _write(p, 30000, np.random.default_rng(9))
    base = loader.read_csv_native(p, num_threads=1)
    for k in (2, 3, 7, 64):
        got = loader.read_csv_native(p, num_threads=k)
        assert all(np.array_equal(a, b) for a, b in zip(base, got)), k


@pytest.mark.parametrize("bad_line", ['01/02/1998,"09:30",1,1,1,1,5',            # quotes
                                      "Date,Time,Open,High,Low,Close,Volume",      # header
                                      "01/02/1998,09:30,1,1,1,,5",                 # missing value
                                      "01/02/1998,09:30,1,1,1,1",                  # 6 fields
                                      "01/02/1998,09:30,1,1,1,1,5,6",              # 8 fields
                                      "01/02/1998,9:30 AM,1,1,1,1,5",              # other time format
                                      "01/02/1998,09:30,1,1,nan,1,5"])             # non-numeric price
def test_unsupported_syntax_is_reported_not_guessed(tmp_path, bad_line):
    p = str(tmp_path / "m.csv")
    lines = _write(p, 50, np.random.default_rng(3))
    with open(p, "w") as f:
        f.write("\n".join(lines[:20] + [bad_line] + lines[20:]) + "\n")
    assert loader.read_csv_native(p) is None
    with pytest.raises(Exception, match="native reader"):
        loader.read_market_csv(p, 4, reader="native")


def test_auto_reader_falls_back_to_pandas_for_such_files(tmp_path):
    rng = np.random.default_rng(5)
    rows = [f'01/{1 + i // 30:02d}/2001,"{9 + (30 + i % 30) // 60:02d}:{(30 + i % 30) % 60:02d}",{10 + i * 0.01:.2f},11,9,10.5,100'
            for i in range(120)]
    p = str(tmp_path / "q.csv")
    with open(p, "w") as f:
        f.write("\n".join(rows) + "\n")
    host = loader.read_market_csv(p, 5)                       # quoted times: pandas path
    ref = loader.read_market_csv(p, 5, reader="pandas")
    assert np.array_equal(host.prices, ref.prices) and np.array_equal(host.seg_start, ref.seg_start)
    assert host.prices.shape == (120, 4) and len(host.seg_start) == 3
    del rng


def test_missing_and_empty_files(tmp_path):
    from finenvs_b200 import _lib

    with pytest.raises(_lib.FeError, match="cannot open"):
        loader.read_csv_native(str(tmp_path / "nope.csv"))
    p = str(tmp_path / "empty.csv")
    open(p, "w").close()
    d, s, o = loader.read_csv_native(p)
    assert len(d) == 0 and o.shape == (0, 4)
    with pytest.raises(Exception, match="no trading day"):
        loader.read_market_csv(p, 4)


@pytest.mark.parametrize("reader", ["native", "pandas"])
def test_both_readers_build_the_same_tables(tmp_path, reader):
    p = str(tmp_path / "m.csv")
    _write(p, 5000, np.random.default_rng(11))
    a = loader.read_market_csv(p, 7, reader=reader)
    b = loader.read_market_csv(p, 7)
    assert np.array_equal(a.prices.view(np.int64), b.prices.view(np.int64))
    assert np.array_equal(a.seg_start, b.seg_start) and np.array_equal(a.seg_len_raw, b.seg_len_raw)
    assert os.path.getsize(p) > 0
