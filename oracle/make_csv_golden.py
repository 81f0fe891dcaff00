"""Generates tests/golden/ref_csv.pkl.gz by running the UNMODIFIED reference's ``read_csv_block``
(nums/core/systems/filesystem.py:157-212) and ``Batch`` (storage/utils.py:23-62) on small synthetic
files.  Build container only (needs /root/reference):  python -m oracle.make_csv_golden

Each case stores the file's bytes, the call arguments and either the returned array or the name of
the exception the reference raised.
"""
import gzip
import os
import pickle
import tempfile

import numpy as np

from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def files():
    rng = np.random.default_rng(2024)
    out = {}
    x = rng.standard_normal((60, 29))
    x[:, 0] = rng.integers(0, 2, 60)
    out["higgs_like"] = "".join(",".join("%.18e" % v for v in row) + "\n" for row in x).encode()
    y = rng.uniform(-100, 100, (41, 5))
    out["header_short"] = ("a,b,c,d,e\n" + "".join(",".join("%.6f" % v for v in row) + "\n" for row in y)).encode()
    out["crlf"] = "".join(",".join("%.17g" % v for v in row) + "\r\n" for row in y[:17]).encode()
    out["no_trailing_newline"] = ("".join(";".join("%.10g" % v for v in row) + "\n" for row in y[:9])).encode()[:-1]
    out["ints"] = "".join(",".join(str(int(v)) for v in row) + "\n" for row in rng.integers(-10**12, 10**12, (23, 4))).encode()
    out["small_ints"] = "".join(",".join("%.3f" % v for v in row) + "\n" for row in rng.uniform(-50, 50, (19, 3))).encode()
    out["bools"] = "".join(",".join(str(int(v)) for v in row) + "\n" for row in rng.integers(0, 2, (15, 6))).encode()
    out["specials"] = (b"nan,inf,-inf,+5, 7.5 ,1e-320\n1e400,-0.0,.5,5.,1E5,1_0.5\n"
                       b"4.9406564584124654e-324,2.4703282292062327e-324,9007199254740993,0.1,123456789012345678901234567890,1e23\n")
    out["tabs"] = "".join("\t".join("%.12e" % v for v in row) + "\n" for row in y[:11]).encode()
    out["ragged"] = b"1,2,3\n4,5\n6,7,8\n"
    out["bad_literal"] = b"1,2,3\n4,abc,6\n"
    out["empty_field"] = b"1,,3\n4,5,6\n"
    out["blank_line"] = b"1,2\n\n3,4\n"
    return out


def main():
    ref_loader.load()
    from nums.core.systems import filesystem
    from nums.core.storage import utils as storage_utils
    cases = []
    tmp = tempfile.mkdtemp()
    blobs = files()

    def run(name, start, end, dtype, delimiter, header):
        path = os.path.join(tmp, name + ".csv")
        try:
            arr, shape = filesystem.read_csv_block(path, start, end, dtype, delimiter, header)
            result = ("ok", np.asarray(arr), tuple(shape))
        except Exception as exc:  # noqa: BLE001
            result = ("raises", type(exc).__name__, None)
        cases.append({"file": name, "start": start, "end": end, "dtype": np.dtype(dtype).name, "delimiter": delimiter,
                      "header": header, "result": result})

    for name, blob in blobs.items():
        with open(os.path.join(tmp, name + ".csv"), "wb") as f:
            f.write(blob)
    spec = {"higgs_like": (np.float64, ",", False), "header_short": (np.float64, ",", True), "crlf": (np.float64, ",", False),
            "no_trailing_newline": (np.float64, ";", False), "ints": (np.int64, ",", False),
            "small_ints": (np.int32, ",", False), "bools": (np.bool_, ",", False), "specials": (np.float64, ",", False),
            "tabs": (np.float32, "\t", False), "ragged": (np.float64, ",", False), "bad_literal": (np.float64, ",", False),
            "empty_field": (np.float64, ",", False), "blank_line": (np.float64, ",", False)}
    batch_table = {}
    for name, (dtype, delimiter, header) in spec.items():
        size = len(blobs[name])
        for workers in (1, 3, 4, 7):
            chunks = storage_utils.Batch.from_num_batches(size, workers).batches
            batch_table[(size, workers)] = [list(c) for c in chunks]
            for start, end in chunks:
                run(name, int(start), int(end), dtype, delimiter, header)
    # every (start, end) on a fine grid of a small file with a trailing newline: pins the ownership rule,
    # including lines that start exactly on a chunk boundary
    size = len(blobs["bools"])
    for start in range(0, size, 5):
        for end in range(start + 1, size + 1, 7):
            run("bools", start, end, np.bool_, ",", False)
    size = len(blobs["crlf"])
    for start in range(0, size, 37):
        for end in range(start + 1, size + 1, 53):
            run("crlf", start, end, np.float64, ",", False)
    out = os.path.join(ROOT, "tests", "golden", "ref_csv.pkl.gz")
    with gzip.open(out, "wb") as f:
        pickle.dump({"files": blobs, "cases": cases, "batches": batch_table}, f, protocol=4)
    ok = sum(1 for c in cases if c["result"][0] == "ok")
    print("%s: %d cases (%d ok, %d raising), %d bytes" % (out, len(cases), ok, len(cases) - ok, os.path.getsize(out)))


if __name__ == "__main__":
    main()
