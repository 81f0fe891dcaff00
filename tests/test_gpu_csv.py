"""Device CSV ingest (nums_csv_index / nums_csv_parse behind cuda_compute.read_csv_block) against the
golden results of the reference's read_csv_block and against the oracle on larger files: bit-exact
values, same shapes, same exceptions."""
import os

import numpy as np
import pytest

from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def _same(got, want):
    return got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want, equal_nan=got.dtype.kind == "f") \
        and (got.dtype.kind != "f" or np.array_equal(np.signbit(got), np.signbit(want)))


def test_read_csv_block_matches_reference_golden(cuda_system, tmp_path):
    from nums_b200 import cuda_compute as cc
    gold = load_golden("ref_csv.pkl.gz")
    for name, blob in gold["files"].items():
        (tmp_path / (name + ".csv")).write_bytes(blob)
    checked = 0
    for case in gold["cases"]:
        path = str(tmp_path / (case["file"] + ".csv"))
        dtype = np.dtype(case["dtype"]).type
        kind, payload, shape = case["result"]
        if kind == "raises":
            with pytest.raises(ValueError):
                cc.read_csv_block(path, case["start"], case["end"], dtype, case["delimiter"], case["header"])
            assert payload == "ValueError"
        else:
            block, got_shape = cc.read_csv_block(path, case["start"], case["end"], dtype, case["delimiter"], case["header"])
            assert tuple(got_shape) == shape, case
            assert _same(cuda_system.get(block), payload), case
        checked += 1
    assert checked == len(gold["cases"]) >= 1400


@pytest.mark.parametrize("fmt,cols,rows", [("%.18e", 29, 20000), ("%.6f", 8, 50000), ("%.17g", 3, 30001)])
def test_read_csv_large_file_bit_exact(cuda_system, tmp_path, fmt, cols, rows):
    """Multi-megabyte files (many 8 KiB tiles, fields straddling tile and chunk boundaries) through the
    FileSystem.read_csv mirror on the device, against the oracle chunk by chunk."""
    from nums_b200 import blocks
    from oracle import csv_oracle
    rng = np.random.default_rng(rows)
    x = rng.standard_normal((rows, cols)) * 10.0 ** rng.integers(-8, 9, (rows, 1))
    path = str(tmp_path / "big.csv")
    np.savetxt(path, x, delimiter=",", fmt=fmt)
    size = os.path.getsize(path)
    fs = blocks.FileSystem(cuda_system)
    parts = fs.read_csv(path, np.float64, ",", False, num_workers=4)
    want = [csv_oracle.read_csv_block(path, s, e, np.float64, ",", False)[0] for s, e in csv_oracle.batches(size, 4)]
    want = [w for w in want if w.shape[0] > 0]
    assert len(parts) == len(want)
    for got, ref in zip(parts, want):
        assert _same(got.get(), ref)
    if fmt != "%.6f":       # round-trip formats: the parsed values are the generated doubles themselves
        whole = cuda_system.get(fs.system.call("read_csv_block", path, 0, size, np.float64, ",", False, syskwargs={})[0])
        assert np.array_equal(whole, x)


def test_read_csv_unsupported_inputs_fail_loudly(cuda_system, tmp_path):
    from nums_b200 import cuda_compute as cc
    p = tmp_path / "hex.csv"
    p.write_bytes(b"0x1p3,2\n3,4\n")
    with pytest.raises(NotImplementedError):
        cc.read_csv_block(str(p), 0, 12, np.float64, ",", False)
    with pytest.raises(NotImplementedError):
        cc.read_csv_block(str(p), 0, 12, int, ",", False)


def test_block_persistence_round_trip_on_device(cuda_system, tmp_path):
    """write_fs / read_fs / delete_fs with device blocks: same files as the reference's format
    (one pickled ndarray per block + meta.pkl), bit-exact round trip for every dtype."""
    import pickle
    from nums_b200 import blocks
    fs = blocks.FileSystem(cuda_system)
    app = blocks.ArrayApp(cuda_system)
    rng = np.random.default_rng(12)
    for k, dt in enumerate((np.float64, np.float32, np.int64, np.int32, np.bool_)):
        x = (rng.standard_normal((45, 33)) * 50).astype(dt)
        X = app.array(x, (20, 16))
        target = str(tmp_path / ("arr%d" % k))
        fs.write_fs(X, target)
        with open(os.path.join(target, "1_2.pkl"), "rb") as fh:
            on_disk = pickle.load(fh)
        assert isinstance(on_disk, np.ndarray) and on_disk.dtype == x.dtype and np.array_equal(on_disk, x[20:40, 32:33])
        back = fs.read_fs(target)
        got = back.get()
        assert got.dtype == x.dtype and np.array_equal(got, x)
        assert np.array_equal((back + back).get(), x + x)        # the loaded blocks are ordinary device blocks
        fs.delete_fs(target)
        assert os.listdir(target) == []
