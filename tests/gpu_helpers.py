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


def compress_batch_host_packed(units):
    """sb_compress_batch_host_packed: the library lays the streams out back to back and reports the offsets."""
    s = snap()
    L = lib()
    n = len(units)
    lens = np.array([len(u) for u in units], dtype=np.uint32)
    in_offs = np.zeros(n, dtype=np.uint64)
    if n:
        in_offs[1:] = np.cumsum(lens[:-1].astype(np.uint64))
    inbuf = np.frombuffer(b"".join(units) + b"\0", dtype=np.uint8).copy()
    cap = int(sum(L.sb_max_compress_len(int(x)) for x in lens)) + 16
    out = np.zeros(cap, dtype=np.uint8)
    out_offs = np.zeros(n + 1, dtype=np.uint64)
    out_lens = np.zeros(max(n, 1), dtype=np.uint32)
    e = s._lib.SbError()
    rc = L.sb_compress_batch_host_packed(inbuf.ctypes.data, in_offs.ctypes.data, lens.ctypes.data, out.ctypes.data, cap,
                                         out_offs.ctypes.data, out_lens.ctypes.data, n, C.byref(e))
    if rc:
        raise s.error.from_c(e)
    streams = [bytes(out[int(out_offs[i]):int(out_offs[i]) + int(out_lens[i])]) for i in range(n)]
    dense = all(int(out_offs[i + 1]) == int(out_offs[i]) + int(out_lens[i]) for i in range(n))
    return streams, dense, int(out_offs[n])


def frame_encode_device_ws(data, ident=True, want_index=True):
    """sb_frame_encode_device_ws over a torch device copy of `data`; returns (stream bytes, chunk offsets, result)."""
    import torch
    s = snap()
    L = lib()
    dev = torch.device("cuda:0")
    n = len(data)
    t_in = torch.frombuffer(bytearray(data) + bytearray(16), dtype=torch.uint8).to(dev)
    cap = L.sb_frame_max_len(n)
    t_out = torch.zeros(cap + 16, dtype=torch.uint8, device=dev)
    nchunks = (n + 65535) // 65536
    t_offs = torch.zeros(nchunks + 1, dtype=torch.int64, device=dev)
    t_res = torch.zeros(64, dtype=torch.uint8, device=dev)
    sb = L.sb_frame_encode_scratch_bytes(n)
    t_scr = torch.empty(sb + 256, dtype=torch.uint8, device=dev)
    e = s._lib.SbError()
    st = torch.cuda.current_stream().cuda_stream
    rc = L.sb_frame_encode_device_ws(t_in.data_ptr(), n, t_out.data_ptr(), cap, 1 if ident else 0,
                                     t_offs.data_ptr() if want_index else None, t_res.data_ptr(), t_scr.data_ptr(), sb + 256, st,
                                     C.byref(e))
    if rc:
        raise s.error.from_c(e)
    torch.cuda.synchronize()
    res = s._lib.SbFrameResult.from_buffer_copy(bytes(t_res.cpu().numpy()[:C.sizeof(s._lib.SbFrameResult)]))
    stream = bytes(t_out[:res.bytes].cpu().numpy())
    return stream, [int(x) for x in t_offs.cpu().numpy()], res


def frame_decode_device(stream, cap, index=None, fragment=False, ws=False):
    """sb_frame_decode_device(_ws) over a device copy of `stream`; returns (status tuple, produced bytes)."""
    import torch
    s = snap()
    L = lib()
    dev = torch.device("cuda:0")
    n = len(stream)
    t_in = torch.frombuffer(bytearray(stream) + bytearray(16), dtype=torch.uint8).to(dev)
    t_out = torch.full((cap + 16,), 0xEE, dtype=torch.uint8, device=dev)
    t_idx = torch.tensor(index, dtype=torch.int64, device=dev) if index is not None else None
    nidx = len(index) - 1 if index is not None else 0
    e = s._lib.SbError()
    st = torch.cuda.current_stream().cuda_stream
    if ws:
        maxc = max(nidx + 1, n // 8 + 16)
        sb = L.sb_frame_decode_scratch_bytes(maxc)
        t_scr = torch.empty(sb + 256, dtype=torch.uint8, device=dev)
        t_res = torch.zeros(64, dtype=torch.uint8, device=dev)
        rc = L.sb_frame_decode_device_ws(t_in.data_ptr(), n, t_out.data_ptr(), cap, t_idx.data_ptr() if t_idx is not None else None,
                                         nidx, 1 if fragment else 0, t_res.data_ptr(), t_scr.data_ptr(), sb + 256, maxc, st, C.byref(e))
        if rc:
            raise s.error.from_c(e)
        torch.cuda.synchronize()
        res = s._lib.SbFrameResult.from_buffer_copy(bytes(t_res.cpu().numpy()[:C.sizeof(s._lib.SbFrameResult)]))
    else:
        res = s._lib.SbFrameResult()
        rc = L.sb_frame_decode_device(t_in.data_ptr(), n, t_out.data_ptr(), cap, t_idx.data_ptr() if t_idx is not None else None,
                                      nidx, 1 if fragment else 0, C.byref(res), st, C.byref(e))
        if rc:
            raise s.error.from_c(e)
    out = bytes(t_out[:res.bytes].cpu().numpy())
    guard = bytes(t_out[cap:cap + 16].cpu().numpy())
    assert guard == b"\xee" * 16
    status = ("Ok", 0, 0, 0) if res.status.code == 0 else err_tuple(s.error.from_c(res.status))
    return status, out
