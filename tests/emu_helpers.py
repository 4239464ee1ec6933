"""ctypes access to the CPU warp-emulator build of the kernel bodies (TEST TOOLING)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "emu", "_build", "libemu_kernels.so")


class SbError(C.Structure):
    _fields_ = [("code", C.c_uint32), ("_pad", C.c_uint32), ("a", C.c_uint64), ("b", C.c_uint64), ("c", C.c_uint64)]


class SbBatch(C.Structure):
    _fields_ = [
        ("in_ptrs", C.c_void_p), ("in_base", C.c_void_p), ("in_stride", C.c_uint64),
        ("in_lens", C.c_void_p), ("in_len_uniform", C.c_uint32),
        ("out_ptrs", C.c_void_p), ("out_base", C.c_void_p), ("out_stride", C.c_uint64),
        ("out_caps", C.c_void_p), ("out_cap_uniform", C.c_uint32),
        ("out_lens", C.c_void_p), ("statuses", C.c_void_p), ("count", C.c_uint32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(HERE, "emu", f) for f in ("emu_kernels.cpp", "simt_emu.cpp", "simt_emu.h")]
        csrc = os.path.join(os.path.dirname(HERE), "rust-snappy_b200", "csrc")
        srcs += [os.path.join(csrc, f) for f in os.listdir(csrc)]
        if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(s) for s in srcs):
            subprocess.check_call([os.path.join(HERE, "emu", "build.sh")])
        _lib = C.CDLL(SO)
    return _lib


ERR = {0: "Ok", 1: "TooBig", 2: "BufferTooSmall", 3: "Empty", 4: "Header", 5: "HeaderMismatch", 6: "Literal",
       7: "CopyRead", 8: "CopyWrite", 9: "Offset"}


def _pack(chunks, slack=64):
    """Concatenate byte strings into one numpy buffer with per-unit offsets."""
    offs, at = [], 0
    for c in chunks:
        offs.append(at)
        at += len(c) + slack
    buf = np.zeros(max(at, 1), dtype=np.uint8)
    for o, c in zip(offs, chunks):
        buf[o:o + len(c)] = np.frombuffer(c, dtype=np.uint8)
    return buf, offs


def compress_units(units, with_header=True, grid=1, hybrid=False, chains=0, out_cap=None, statuses=False, crcs=False):
    """Run the K1 kernel body under the emulator over independent units (<=64KB each).
    hybrid: 7 shared-memory-table chains + 4 chains with tables in global memory (else 7 + 0)."""
    inbuf, inoffs = _pack(units, slack=3)
    stride = 76544
    out = np.full(stride * len(units) + 64, 0xEE, dtype=np.uint8)
    lens = np.array([len(u) for u in units], dtype=np.uint32)
    in_ptrs = np.array([inbuf.ctypes.data + o for o in inoffs], dtype=np.uint64)
    out_lens = np.zeros(len(units), dtype=np.uint32)
    st = (SbError * len(units))()
    b = SbBatch()
    b.in_ptrs = in_ptrs.ctypes.data
    b.in_lens = lens.ctypes.data
    b.out_base = out.ctypes.data
    b.out_stride = stride
    b.out_cap_uniform = stride if out_cap is None else out_cap
    b.out_lens = out_lens.ctypes.data
    if statuses:
        b.statuses = C.addressof(st)
    b.count = len(units)
    crc = np.zeros(len(units), dtype=np.uint32)
    lib().emu_compress_batch(C.byref(b), (1 if with_header else 0) | (0x400 if hybrid else 0) | (chains << 20), grid,
                             C.c_void_p(crc.ctypes.data if crcs else None))
    res = [bytes(out[i * stride:i * stride + int(out_lens[i])]) for i in range(len(units))]
    if crcs:
        return res, [int(x) for x in crc]
    if statuses:
        return res, [(ERR.get(e.code, str(e.code)), e.a, e.b) for e in st]
    return res


def decompress_units(streams, caps, grid=1, block=32):
    """Run the K2 kernel body under the emulator. Returns [(status tuple, bytes)]."""
    inbuf, inoffs = _pack(streams, slack=0)
    ocap = [max(c, 0) for c in caps]
    ooffs, at = [], 0
    for c in ocap:
        ooffs.append(at)
        at += c + 16
    out = np.full(at + 16, 0xEE, dtype=np.uint8)
    lens = np.array([len(s) for s in streams], dtype=np.uint32)
    in_ptrs = np.array([inbuf.ctypes.data + o for o in inoffs], dtype=np.uint64)
    out_ptrs = np.array([out.ctypes.data + o for o in ooffs], dtype=np.uint64)
    out_caps = np.array(ocap, dtype=np.uint32)
    out_lens = np.zeros(len(streams), dtype=np.uint32)
    st = (SbError * len(streams))()
    b = SbBatch()
    b.in_ptrs = in_ptrs.ctypes.data
    b.in_lens = lens.ctypes.data
    b.out_ptrs = out_ptrs.ctypes.data
    b.out_caps = out_caps.ctypes.data
    b.out_lens = out_lens.ctypes.data
    b.statuses = C.addressof(st)
    b.count = len(streams)
    lib().emu_decompress_batch(C.byref(b), grid, block)
    res = []
    for i in range(len(streams)):
        e = st[i]
        res.append(((ERR.get(e.code, str(e.code)), e.a, e.b, e.c),
                    bytes(out[ooffs[i]:ooffs[i] + int(out_lens[i])]),
                    bytes(out[ooffs[i] + ocap[i]:ooffs[i] + ocap[i] + 16])))
    return res


class SbFrameResult(C.Structure):
    _fields_ = [("status", SbError), ("bytes", C.c_uint64), ("nchunks", C.c_uint32), ("_pad", C.c_uint32)]


def frame_encode(data, ident=True):
    """K4 assembly around K1 (with the fused chunk checksum) under the emulator: (stream, chunk offsets, result)."""
    n = len(data)
    src = np.frombuffer(bytes(data) + b"\0" * 16, dtype=np.uint8).copy()
    nchunks = (n + 65535) // 65536
    cap = 10 + nchunks * (8 + 76490)
    out = np.full(cap + 16, 0xEE, dtype=np.uint8)
    offs = np.zeros(nchunks + 2, dtype=np.uint64)
    res = SbFrameResult()
    lib().emu_frame_encode(C.c_void_p(src.ctypes.data), C.c_uint64(n), C.c_void_p(out.ctypes.data), C.c_uint64(cap), 1 if ident else 0,
                           C.c_void_p(offs.ctypes.data), C.byref(res))
    return bytes(out[:res.bytes]), [int(x) for x in offs[:nchunks + 1]], res


def frame_decode(stream, cap, index=None, fragment=False, max_chunks=None):
    """K5 (parse or walk, scan, decode + checksum, finish) under the emulator: (status tuple, produced bytes)."""
    n = len(stream)
    src = np.frombuffer(bytes(stream) + b"\0" * 16, dtype=np.uint8).copy()
    out = np.full(cap + 16, 0xEE, dtype=np.uint8)
    res = SbFrameResult()
    idx = np.array(index, dtype=np.uint64) if index is not None else None
    nidx = len(index) - 1 if index is not None else 0
    maxc = max_chunks if max_chunks is not None else max(nidx + 1, n // 8 + 16)
    lib().emu_frame_decode(C.c_void_p(src.ctypes.data), C.c_uint64(n), C.c_void_p(out.ctypes.data), C.c_uint64(cap),
                           C.c_void_p(idx.ctypes.data) if idx is not None else None, nidx, 1 if fragment else 0, C.byref(res), maxc)
    assert bytes(out[cap:cap + 16]) == b"\xee" * 16
    e = res.status
    status = (ERR.get(e.code, {10: "StreamHeader", 11: "StreamHeaderMismatch", 12: "UnsupportedChunkType", 13: "UnsupportedChunkLength",
                                14: "Checksum", 100: "UnexpectedEof", 202: "Invalid"}.get(e.code, str(e.code))), e.a, e.b, e.c)
    return status, bytes(out[:res.bytes]), res
