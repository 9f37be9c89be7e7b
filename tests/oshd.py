"""Reader for the flat "OSHD1" record files written by oracle/ref_driver.cpp and by
the package's own dump helper. Test infrastructure."""
import gzip
import struct

import numpy as np

_DT = {0: np.int8, 1: np.int32, 2: np.int64, 3: np.float64}


def read_oshd(path):
    out = {}
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        magic = f.read(6)
        assert magic == b"OSHD1\n", magic
        while True:
            hdr = f.read(4)
            if len(hdr) < 4:
                break
            (nl,) = struct.unpack("<I", hdr)
            name = f.read(nl).decode()
            (dt,) = struct.unpack("<B", f.read(1))
            (cnt,) = struct.unpack("<Q", f.read(8))
            dtype = np.dtype(_DT[dt])
            data = np.frombuffer(f.read(cnt * dtype.itemsize), dtype=dtype).copy()
            out[name] = data
    return out


def write_oshd(path, arrays):
    inv = {np.dtype(v): k for k, v in _DT.items()}
    with open(path, "wb") as f:
        f.write(b"OSHD1\n")
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            f.write(struct.pack("<B", inv[a.dtype]))
            f.write(struct.pack("<Q", a.size))
            f.write(a.tobytes())
