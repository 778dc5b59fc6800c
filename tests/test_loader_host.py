"""Host half of the loader (CSV -> flat prices + segment table) against what the reference built
from the same CSVs (golden fixtures), plus the reference's file-lookup errors.  CPU only."""
import os

import numpy as np
import pytest

from parity_utils import GOLDEN, load_trace


def _write_dummy_csv(tmp_path, name):
    z = load_trace("dummy_csv.npz")
    d = tmp_path / f"data_{name}"
    d.mkdir()
    p = d / "dummy.csv"
    with open(p, "w") as f:
        for date, time, ohlc, vol in zip(z[f"{name}_csv_date"], z[f"{name}_csv_time"], z[f"{name}_csv_ohlc"],
                                         z[f"{name}_csv_volume"]):
            f.write(f"{date.decode()},{time.decode()},{float(ohlc[0])!r},{float(ohlc[1])!r},{float(ohlc[2])!r},{float(ohlc[3])!r},{vol}\n")
    return str(d), str(p)


@pytest.mark.parametrize("trace", ["kat_ibm_w390.npz", "trace_ibm_w60.npz", "trace_ibm_w4.npz", "trace_oih_w390.npz",
                                   "trace_oih_w60.npz", "trace_oih_w4.npz", "trace_spy_w60.npz", "trace_spy_w390.npz"])
def test_flat_tables_equal_the_references(tmp_path, trace):
    from finenvs_b200.data import loader

    z = load_trace(trace)
    name = trace.split("_")[1].upper()
    _, path = _write_dummy_csv(tmp_path, name)
    host = loader.read_market_csv(path, int(z["window"]))
    assert np.array_equal(host.prices, z["prices"])            # between_time + column order
    assert np.array_equal(host.seg_start, z["seg_start"])      # :141-152 incl. skipped first days
    assert np.array_equal(host.seg_len_raw, z["seg_len_raw"])
    # shape of the reference's padded tensors is implied by the table
    assert tuple(z["pe_shape"]) == (len(host.seg_start), int(host.seg_len_raw.max()), 4)


def test_padded_reconstruction_digest(tmp_path):
    """Row j of the reference's price_environments[d] is row seg_start[d]+j of the flat series."""
    import hashlib

    z = load_trace("trace_oih_w60.npz")
    D, L = len(z["seg_start"]), int(z["seg_len_raw"].max())
    pe = np.full((D, L, 4), np.nan)
    le = np.full((D, L, 4), np.nan)
    for d in range(D):
        s, n = int(z["seg_start"][d]), int(z["seg_len_raw"][d])
        pe[d, :n] = z["prices"][s:s + n]
        le[d, :n] = z["logret"][s:s + n]
    dg = lambda a: hashlib.sha256(np.ascontiguousarray(np.nan_to_num(a, nan=-12345.0)).tobytes()).hexdigest()
    assert dg(pe) == str(z["pe_digest"]) and dg(le) == str(z["le_digest"])


def test_file_lookup_errors_mirror_the_reference(tmp_path):
    from finenvs_b200.data import loader

    with pytest.raises(Exception, match="dataset_key expected"):
        loader.determine_file_key("nonsense")
    assert loader.determine_file_key("cross_validation") == "valid"
    d = tmp_path / "data_x"
    d.mkdir()
    with pytest.raises(Exception, match="No file was found"):
        loader.find_file_by_key(str(d), "train")
    (d / "a_train.csv").write_text("")
    (d / "b_train.csv").write_text("")
    with pytest.raises(Exception, match="More than one file"):
        loader.find_file_by_key(str(d), "train")
    assert loader.get_data_dir_name(str(d)) == str(d)            # contains "data" -> verbatim (:47-51)
    assert loader.get_data_dir_name("IBM").endswith(os.path.join("data", "IBM"))


def test_segment_table_edge_cases():
    from finenvs_b200.data import loader

    dates = np.array(["d0"] * 3 + ["d1"] * 5 + ["d2"] * 1 + ["d3"] * 4)
    start, length = loader.segment_table(dates, 4)
    # d0 (start -4) and d1 (start -1) are skipped (:134); d2 has a single bar
    assert start.tolist() == [4, 5] and length.tolist() == [5, 8]
    start, length = loader.segment_table(dates, 1)
    assert start.tolist() == [2, 7, 8] and length.tolist() == [6, 2, 5]
    # integer keys (the native reader's date hashes): same tables, also when a date comes back after another one
    keys = np.array([7] * 3 + [-2] * 5 + [99] * 1 + [4] * 4, np.int64)
    assert [a.tolist() for a in loader.segment_table(keys, 4)] == [[4, 5], [5, 8]]
    back = np.array(["a", "a", "b", "b", "b", "a", "c", "c"])
    codes = np.array([hash(x) for x in back], np.int64)
    for w in (1, 2):
        assert [a.tolist() for a in loader.segment_table(codes, w)] == [a.tolist() for a in loader.segment_table(back, w)]
    assert [a.tolist() for a in loader.segment_table(np.zeros(0, np.int64), 3)] == [[], []]
    s, n = loader.regular_segments(1000, 100, 60)
    assert s.tolist() == [40 + 100 * i for i in range(9)] and (n == 160).all()
