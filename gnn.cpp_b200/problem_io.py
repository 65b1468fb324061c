"""Binary problem (.gcnp) / result (.gcno) files exchanged with the C++ drivers (host/main.cpp,
oracle/ref_driver.cpp).  Little-endian, no alignment padding.

.gcnp: int64 magic 0x47434E50, N, E, L, dims[L+1]; int32 src[E], dst[E]; f32 X[N*F0]; int32 y[N];
       per layer: f32 W[F_l*F_{l-1}], f32 b[F_l]
.gcno: records of (int32 name_len, name, int32 dtype {0:f32,1:i32,2:i64}, int32 ndim, int64 shape[], data)
"""
import struct

import numpy as np

MAGIC = 0x47434E50
_DT = {0: np.float32, 1: np.int32, 2: np.int64}


def write_problem(path, prob):
    cfg = prob.cfg
    L = len(cfg.dims) - 1
    with open(path, "wb") as f:
        f.write(struct.pack("<4q", MAGIC, cfg.N, len(prob.src), L))
        f.write(np.asarray(cfg.dims, dtype=np.int64).tobytes())
        f.write(np.ascontiguousarray(prob.src, dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(prob.dst, dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(prob.X, dtype=np.float32).tobytes())
        f.write(np.ascontiguousarray(prob.y, dtype=np.int32).tobytes())
        for W, b in zip(prob.W, prob.b):
            f.write(np.ascontiguousarray(W, dtype=np.float32).tobytes())
            f.write(np.ascontiguousarray(b, dtype=np.float32).tobytes())


def read_results(path):
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    off = 0
    while off < len(buf):
        (nl,) = struct.unpack_from("<i", buf, off); off += 4
        name = buf[off:off + nl].decode(); off += nl
        dt, nd = struct.unpack_from("<2i", buf, off); off += 8
        shape = struct.unpack_from("<%dq" % nd, buf, off); off += 8 * nd
        n = int(np.prod(shape)) if nd else 1
        dtype = _DT[dt]
        out[name] = np.frombuffer(buf, dtype=dtype, count=n, offset=off).reshape(shape).copy()
        off += n * np.dtype(dtype).itemsize
    return out
