"""Helpers for the -m gpu parity tests: everything goes through the C ABI."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402


def snap():
    return graft.load_package()


def lib():
    return snap()._lib.lib()


def err_tuple(e):
    """Normalise a raised exception to the oracle's (variant, a, b, c) tuple."""
    s = snap()
    if isinstance(e, s.Error):
        return e.as_tuple()
    if isinstance(e, s.UnexpectedEof):
        return ("UnexpectedEof", 0, 0, 0)
    raise e


def compress_batch_host(units):
    """sb_compress_batch_host over independent units (<=64KB each), compact output."""
    s = snap()
    L = lib()
    n = len(units)
    lens = np.array([len(u) for u in units], dtype=np.uint32)
    in_offs = np.zeros(n, dtype=np.uint64)
    if n:
        in_offs[1:] = np.cumsum(lens[:-1].astype(np.uint64))
    inbuf = np.frombuffer(b"".join(units) + b"\0", dtype=np.uint8).copy()
    caps = np.array([L.sb_max_compress_len(int(x)) for x in lens], dtype=np.uint32)
    out_offs = np.zeros(n, dtype=np.uint64)
    # dense destinations: unit k lands right after unit k-1 (we do not know sizes up front,
    # so give every unit its own capacity-sized slot)
    if n:
        out_offs[1:] = np.cumsum(caps[:-1].astype(np.uint64))
    out = np.zeros(int(caps.astype(np.uint64).sum()) + 16, dtype=np.uint8)
    out_lens = np.zeros(n, dtype=np.uint32)
    e = s._lib.SbError()
    rc = L.sb_compress_batch_host(inbuf.ctypes.data, in_offs.ctypes.data, lens.ctypes.data, out.ctypes.data,
                                  out_offs.ctypes.data, caps.ctypes.data, out_lens.ctypes.data, n, C.byref(e))
    if rc:
        raise s.error.from_c(e)
    return [bytes(out[int(o):int(o) + int(k)]) for o, k in zip(out_offs, out_lens)]


def decompress_batch_host(streams, caps):
    s = snap()
    L = lib()
    n = len(streams)
    lens = np.array([len(u) for u in streams], dtype=np.uint32)
    in_offs = np.zeros(n, dtype=np.uint64)
    if n:
        in_offs[1:] = np.cumsum(lens[:-1].astype(np.uint64))
    inbuf = np.frombuffer(b"".join(streams) + b"\0", dtype=np.uint8).copy()
    caps = np.array(caps, dtype=np.uint32)
    out_offs = np.zeros(n, dtype=np.uint64)
    if n:
        out_offs[1:] = np.cumsum(caps[:-1].astype(np.uint64))
    out = np.zeros(int(caps.astype(np.uint64).sum()) + 16, dtype=np.uint8)
    out_lens = np.zeros(n, dtype=np.uint32)
    st = (s._lib.SbError * max(n, 1))()
    e = s._lib.SbError()
    rc = L.sb_decompress_batch_host(inbuf.ctypes.data, in_offs.ctypes.data, lens.ctypes.data, out.ctypes.data,
                                    out_offs.ctypes.data, caps.ctypes.data, out_lens.ctypes.data, C.addressof(st), n,
                                    C.byref(e))
    if rc:
        raise s.error.from_c(e)
    res = []
    for i in range(n):
        code = st[i].code
        if code:
            res.append((s.error.from_c(st[i]).as_tuple(), b""))
        else:
            res.append((("Ok", 0, 0, 0), bytes(out[int(out_offs[i]):int(out_offs[i]) + int(out_lens[i])])))
    return res


def batch_from_tensors(in_t, in_stride, in_len, out_t, out_stride, out_cap, lens_t, status_t, count, in_lens_t=None):
    """sb_batch over torch CUDA tensors with base+stride addressing."""
    s = snap()
    b = s._lib.SbBatch()
    b.in_base = in_t.data_ptr()
    b.in_stride = in_stride
    b.in_len_uniform = in_len
    if in_lens_t is not None:
        b.in_lens = in_lens_t.data_ptr()
    b.out_base = out_t.data_ptr()
    b.out_stride = out_stride
    b.out_cap_uniform = out_cap
    b.out_lens = lens_t.data_ptr()
    if status_t is not None:
        b.statuses = status_t.data_ptr()
    b.count = count
    return b
