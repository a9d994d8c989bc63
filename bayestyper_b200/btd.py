"""Reader/writer for the BTD1 named-array container (oracle dumps, fixtures)."""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

_DT = {0: np.uint8, 1: np.uint16, 2: np.uint32, 3: np.uint64, 4: np.int32, 5: np.float32, 6: np.float64, 7: np.int64}
_DT_INV = {np.dtype(v): k for k, v in _DT.items()}


def read(path) -> dict:
    buf = Path(path).read_bytes()
    assert buf[:4] == b"BTD1", "not a BTD1 file"
    off = 4
    out = {}
    while off < len(buf):
        (nl,) = struct.unpack_from("<I", buf, off); off += 4
        name = buf[off:off + nl].decode(); off += nl
        dt, nd = struct.unpack_from("<BB", buf, off); off += 2
        dims = struct.unpack_from("<" + "Q" * nd, buf, off); off += 8 * nd
        n = int(np.prod(dims)) if nd else 1
        dtype = np.dtype(_DT[dt])
        out[name] = np.frombuffer(buf, dtype, n, off).reshape(dims).copy()
        off += n * dtype.itemsize
    return out


def write(path, arrays: dict) -> None:
    with open(path, "wb") as f:
        f.write(b"BTD1")
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            nb = name.encode()
            f.write(struct.pack("<I", len(nb))); f.write(nb)
            f.write(struct.pack("<BB", _DT_INV[a.dtype], a.ndim))
            f.write(struct.pack("<" + "Q" * a.ndim, *a.shape))
            f.write(a.tobytes())
