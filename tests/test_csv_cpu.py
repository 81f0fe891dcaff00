"""CSV ingest, CPU side: the oracle against the golden results of the real reference, and the number
parser of the kernels (nums_b200/csrc/csv_parse.cuh, compiled for the host) against strtod."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _same(got, want):
    return got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want, equal_nan=got.dtype.kind == "f") \
        and (got.dtype.kind != "f" or np.array_equal(np.signbit(got), np.signbit(want)))


def test_oracle_read_csv_block_matches_reference_golden(tmp_path):
    """All 1433 recorded read_csv_block calls of the reference (results and raised exceptions)."""
    from oracle import csv_oracle
    gold = load_golden("ref_csv.pkl.gz")
    for name, blob in gold["files"].items():
        (tmp_path / (name + ".csv")).write_bytes(blob)
    for case in gold["cases"]:
        path = str(tmp_path / (case["file"] + ".csv"))
        dtype = np.dtype(case["dtype"]).type
        kind, payload, shape = case["result"]
        if kind == "raises":
            with pytest.raises(Exception) as info:
                csv_oracle.read_csv_block(path, case["start"], case["end"], dtype, case["delimiter"], case["header"])
            assert type(info.value).__name__ == payload
        else:
            arr, got_shape = csv_oracle.read_csv_block(path, case["start"], case["end"], dtype, case["delimiter"], case["header"])
            assert tuple(got_shape) == shape, case
            assert _same(arr, payload), case
    for (size, workers), want in gold["batches"].items():
        assert csv_oracle.batches(size, workers) == want
        from nums_b200.blocks import FileSystem
        assert FileSystem.byte_ranges(size, workers) == want


def test_reference_live_when_present(tmp_path):
    """With the reference tree mounted (build container), the oracle equals it on a fresh random file."""
    from oracle import ref_loader, csv_oracle
    if not ref_loader.available():
        pytest.skip("reference tree not present")
    ref_loader.load()
    from nums.core.systems import filesystem
    rng = np.random.default_rng(5)
    path = str(tmp_path / "live.csv")
    with open(path, "w") as f:
        for row in rng.standard_normal((200, 7)):
            f.write(",".join(repr(float(v)) for v in row) + "\n")
    size = os.path.getsize(path)
    for start, end in csv_oracle.batches(size, 5):
        a, sa = filesystem.read_csv_block(path, start, end, np.float64, ",", False)
        b, sb = csv_oracle.read_csv_block(path, start, end, np.float64, ",", False)
        assert tuple(sa) == tuple(sb) and np.array_equal(a, b)


def test_number_parser_host_build_matches_strtod(tmp_path):
    """csv_parse.cuh is __host__ __device__: the same source, built with g++, against glibc's correctly
    rounded strtod on ~1.6 million generated literals plus fixed edge cases."""
    exe = str(tmp_path / "csv_host_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "native", "csv_host_check.cpp")])
    out = subprocess.run([exe, "300000", "7"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    sys.stderr.write(out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 mismatches" in out.stdout


def test_filesystem_read_csv_mirror_over_oracle(tmp_path):
    """FileSystem.read_csv (mirror of filesystem.py:402-439) over the CPU oracle system: the stacked
    blocks equal np.loadtxt except for lines that start exactly on a chunk boundary, which the
    reference's ownership rule drops (pinned by the golden cases above)."""
    from nums_b200 import blocks
    from oracle import csv_oracle
    from tests.helpers import OracleSystem
    rng = np.random.default_rng(6)
    x = rng.standard_normal((257, 6))
    path = str(tmp_path / "m.csv")
    np.savetxt(path, x, delimiter=",")
    fs = blocks.FileSystem(OracleSystem(), csv_oracle.read_csv_block)
    parts = fs.read_csv(path, np.float64, ",", False, num_workers=4)
    got = np.concatenate([p.get() for p in parts], axis=0)
    size = os.path.getsize(path)
    starts = set(np.cumsum([len(line) for line in open(path, "rb")])[:-1].tolist()) | {0}
    dropped = [s for s, _e in csv_oracle.batches(size, 4)[1:] if s in starts]
    assert got.shape[0] == x.shape[0] - len(dropped)
    if not dropped:
        assert np.array_equal(got, x)


def test_read_range_parallel_positional_reads(tmp_path):
    """The host-side file reader of read_csv_block: multi-chunk ranges are read by a thread pool with
    positional reads; content, offsets and the short-read error."""
    from nums_b200 import cuda_compute as cc
    data = np.random.default_rng(8).integers(0, 256, 20_000_003, dtype=np.uint8)
    path = str(tmp_path / "blob.bin")
    data.tofile(path)
    with open(path, "rb") as fh:
        for offset, count in ((0, data.size), (12345, 17_000_000), (data.size - 1003, 1003), (7, 5), (3, 0)):
            buf = np.zeros(count, dtype=np.uint8)
            cc._read_range(fh.fileno(), offset, memoryview(buf), path)
            assert np.array_equal(buf, data[offset:offset + count])
        with pytest.raises(IOError):
            cc._read_range(fh.fileno(), data.size - 10, memoryview(np.zeros(11, dtype=np.uint8)), path)


def test_block_persistence_reference_format(tmp_path):
    """write_fs / read_fs mirror (application.py:154-191) over the CPU oracle system: round trip, the
    reference's directory layout, and -- with the reference tree mounted -- files written by either side
    are read by the other."""
    import pickle
    from nums_b200 import blocks
    from oracle import csv_oracle, ref_loader
    from tests.helpers import OracleSystem

    def np_write(block, filename, entry):
        os.makedirs(filename, exist_ok=True)
        with open(os.path.join(filename, "_".join(map(str, entry)) + ".pkl"), "wb") as fh:
            pickle.dump(np.asarray(block), fh)

    def np_read(filename, entry):
        with open(os.path.join(filename, "_".join(map(str, entry)) + ".pkl"), "rb") as fh:
            return pickle.load(fh)

    def np_delete(filename, entry):
        os.remove(os.path.join(filename, "_".join(map(str, entry)) + ".pkl"))

    system = OracleSystem()
    fs = blocks.FileSystem(system, csv_oracle.read_csv_block, (np_write, np_read, np_delete))
    app = blocks.ArrayApp(system)
    x = np.random.default_rng(11).standard_normal((37, 21))
    X = app.array(x, (16, 8))
    target = str(tmp_path / "arr")
    meta = fs.write_fs(X, target)
    assert sorted(os.listdir(target)) == sorted(["meta.pkl"] + ["%d_%d.pkl" % e for e in X.grid.get_entry_iterator()])
    assert meta["grid_meta"] == {"shape": (37, 21), "block_shape": (16, 8), "dtype": "float64"}
    back = fs.read_fs(target)
    assert back.grid.to_meta() == X.grid.to_meta() and np.array_equal(back.get(), x)
    if ref_loader.available():
        ref_loader.load()
        from nums.core.systems import filesystem as ref_fs
        assert np.array_equal(ref_fs.read_block_fs(target, (1, 2)), x[16:32, 16:21])          # reference reads our files
        ref_fs.write_block_fs(x[:16, :8] * 2.0, target, (0, 0))                                 # we read the reference's
        assert np.array_equal(fs.read_fs(target).get()[:16, :8], x[:16, :8] * 2.0)
        assert ref_fs.read_meta_fs(target)["grid_meta"] == meta["grid_meta"]
    fs.delete_fs(target)
    assert os.listdir(target) == []
