"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's CSV chunk reader.

Follows ``read_csv_block`` (/root/reference/nums/core/systems/filesystem.py:157-212) and the batch
arithmetic of ``FileSystem.read_csv`` (:402-439, ``storage/utils.py:30-62``).  Pinned against the real
reference by ``oracle/make_csv_golden.py`` -> ``tests/golden/ref_csv.pkl.gz`` (``tests/test_oracle.py``).
Only tests, ``smoke()`` and bench.py's CPU legs may import this module.
"""
import numpy as np


def converter(dtype):
    """Field converter per dtype (filesystem.py:160-190, itself adapted from numpy/lib/npyio.py)."""
    if issubclass(dtype, np.bool_):
        return lambda tok: bool(int(tok))                    # :169-170
    if issubclass(dtype, np.uint64):
        return np.uint64                                     # :171-172
    if issubclass(dtype, np.int64):
        return np.int64                                      # :173-174
    if issubclass(dtype, np.integer) or dtype is int:
        return lambda tok: int(float(tok))                   # :175-176
    if issubclass(dtype, np.floating) or dtype is float:
        def as_float(tok):                                   # :163-167
            return float.fromhex(tok) if "0x" in tok else float(tok)
        return as_float
    raise NotImplementedError("oracle read_csv_block: dtype %r" % (dtype,))


def read_csv_block(filename, file_start, file_end, dtype, delimiter, has_header):
    """Rows of the lines that START in [file_start', file_end), file_start' being just past the first
    newline at or after a non-zero file_start (:196-210); header dropped in the first chunk."""
    convert = converter(dtype)
    rows = []
    with open(filename, "r") as fh:
        fh.seek(file_start)
        if file_start != 0:
            while True:                                      # :198-201 (the reference spins at EOF; we stop)
                ch = fh.read(1)
                if ch == "\n" or ch == "":
                    break
        drop_header = has_header and file_start == 0
        while fh.tell() < file_end:                          # :203
            text = fh.readline().strip("\r\n")               # :204
            if drop_header:
                drop_header = False
                continue
            rows.append([convert(tok) for tok in text.split(delimiter)])   # :208-210
    arr = np.array(rows, dtype=dtype)                        # :211
    return arr, arr.shape


def batches(total_size, num_batches):
    """``Batch.from_num_batches(total, n).batches`` (storage/utils.py:30-62)."""
    batch_size = (total_size + num_batches - 1) // num_batches
    if total_size < batch_size:
        return [[0, total_size]]
    starts = list(range(0, total_size, batch_size))
    count = int(total_size / batch_size)
    out = [starts[i:i + 2] for i in range(count)]
    if len(out[-1]) == 1:
        out[-1].append(total_size)
    if out[-1][1] != total_size:
        out.append([out[-1][1], total_size])
    return out
