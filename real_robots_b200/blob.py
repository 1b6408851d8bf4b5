"""Scene asset blob (`*.rrb`) reader / writer.

The blob is the only thing the CUDA library, the CPU oracle and the Python
facade need from the reference's data directory
(`real_robots/data/kuka_gripper_description/{urdf,meshes}` in the reference,
SURVEY.md App. A).  `tools/compile_assets.py` writes it once; everything else
reads it.

Layout (little endian):

    0   char[8]  magic  "RRB200\\0\\0"
    8   u32      version (1)
    12  u32      n_entries
    16  entries[n_entries], 112 bytes each:
          char[48] name (NUL padded)
          u32 dtype code (0=u8 1=i32 2=f32 3=f64 4=i64)
          u32 ndim
          u64 shape[4]
          u64 offset   (from start of file, 64-byte aligned)
          u64 nbytes
          u64 reserved
    ...  data
"""
import struct

import numpy as np

MAGIC = b"RRB200\0\0"
VERSION = 1
_DTYPES = [np.uint8, np.int32, np.float32, np.float64, np.int64]
_ENTRY = struct.Struct("<48sII4QQQQ")
assert _ENTRY.size == 112


def _code(dtype):
    for i, d in enumerate(_DTYPES):
        if np.dtype(dtype) == np.dtype(d):
            return i
    raise TypeError("unsupported dtype %r" % (dtype,))


def write_blob(path, arrays):
    """arrays: dict name -> ndarray (ndim <= 4)."""
    names = list(arrays.keys())
    items = []
    for n in names:
        a = np.ascontiguousarray(arrays[n])
        if a.ndim > 4:
            raise ValueError("%s: ndim > 4" % n)
        if len(n.encode()) > 47:
            raise ValueError("name too long: %s" % n)
        items.append((n, a))
    off = 16 + _ENTRY.size * len(items)
    off = (off + 63) // 64 * 64
    table = []
    for n, a in items:
        shape = list(a.shape) + [0] * (4 - a.ndim)
        table.append(_ENTRY.pack(n.encode(), _code(a.dtype), a.ndim,
                                 *shape, off, a.nbytes, 0))
        off = (off + a.nbytes + 63) // 64 * 64
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<II", VERSION, len(items)))
        for t in table:
            f.write(t)
        for (n, a), t in zip(items, table):
            o = _ENTRY.unpack(t)[7]
            f.seek(o)
            f.write(a.tobytes())
        # pad tail to 64
        end = f.tell()
        pad = (-end) % 64
        if pad:
            f.write(b"\0" * pad)


def read_blob(path):
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:8] != MAGIC:
        raise ValueError("%s: not an RRB200 blob" % path)
    version, n = struct.unpack_from("<II", raw, 8)
    if version != VERSION:
        raise ValueError("%s: blob version %d, expected %d" % (path, version, VERSION))
    out = {}
    for i in range(n):
        e = _ENTRY.unpack_from(raw, 16 + i * _ENTRY.size)
        name = e[0].split(b"\0", 1)[0].decode()
        dtype = _DTYPES[e[1]]
        ndim = e[2]
        shape = tuple(int(s) for s in e[3:3 + ndim])
        off, nbytes = e[7], e[8]
        a = np.frombuffer(raw, dtype=dtype, count=nbytes // np.dtype(dtype).itemsize,
                          offset=off).reshape(shape)
        out[name] = a
    return out
