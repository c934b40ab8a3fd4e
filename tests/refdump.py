"""Reader for the named-array dump files written by oracle/ref_harness/ref_driver.cu
(`name dtype count\\n` header line followed by raw little-endian data)."""
import numpy as np

_DT = {"f32": np.float32, "i32": np.int32, "f64": np.float64}


def load(path):
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        nl = data.index(b"\n", pos)
        name, dt, cnt = data[pos:nl].decode().split()
        dt = _DT[dt]
        nbytes = int(cnt) * np.dtype(dt).itemsize
        out[name] = np.frombuffer(data, dtype=dt, count=int(cnt), offset=nl + 1).copy()
        pos = nl + 1 + nbytes
    return out
