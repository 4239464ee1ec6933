"""ctypes binding of the CPU oracle (oracle/snappy_oracle.c).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs. The product package
(rust-snappy_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

ERROR_NAMES = {
    0: "Ok", 1: "TooBig", 2: "BufferTooSmall", 3: "Empty", 4: "Header",
    5: "HeaderMismatch", 6: "Literal", 7: "CopyRead", 8: "CopyWrite",
    9: "Offset", 10: "StreamHeader", 11: "StreamHeaderMismatch",
    12: "UnsupportedChunkType", 13: "UnsupportedChunkLength", 14: "Checksum",
    100: "UnexpectedEof",
}


class OrcError(C.Structure):
    _fields_ = [("code", C.c_uint32), ("_pad", C.c_uint32),
                ("a", C.c_uint64), ("b", C.c_uint64), ("c", C.c_uint64)]

    def tuple(self):
        return (ERROR_NAMES.get(self.code, str(self.code)), self.a, self.b, self.c)


def build(force=False):
    src = os.path.join(_HERE, "snappy_oracle.c")
    hdr = os.path.join(_HERE, "snappy_oracle.h")
    stale = (not os.path.exists(_SO)
             or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, sz, szp, ep = C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(OrcError)
        L.orc_max_compress_len.restype = sz
        L.orc_max_compress_len.argtypes = [sz]
        L.orc_compress.argtypes = [u8p, sz, C.c_void_p, sz, szp, ep]
        L.orc_decompress_len.argtypes = [u8p, sz, szp, ep]
        L.orc_decompress.argtypes = [u8p, sz, C.c_void_p, sz, szp, ep]
        L.orc_crc32c.restype = C.c_uint32
        L.orc_crc32c.argtypes = [u8p, sz]
        L.orc_crc32c_bitwise.restype = C.c_uint32
        L.orc_crc32c_bitwise.argtypes = [u8p, sz]
        L.orc_crc32c_masked.restype = C.c_uint32
        L.orc_crc32c_masked.argtypes = [u8p, sz]
        L.orc_compress_frame.argtypes = [u8p, sz, C.c_void_p, szp]
        L.orc_frame_max_len.restype = sz
        L.orc_frame_max_len.argtypes = [sz]
        L.orc_frame_encode.argtypes = [u8p, sz, C.c_void_p, sz, szp]
        L.orc_frame_decode.argtypes = [u8p, sz, C.c_void_p, sz, szp, ep]
        L.orc_bench_compress_mt.restype = C.c_double
        L.orc_bench_compress_mt.argtypes = [u8p, sz, sz, C.c_uint64, C.c_uint64, C.c_uint64,
                                            C.c_int, C.POINTER(C.c_uint64)]
        L.orc_fingerprint_blocks_mt.restype = C.c_double
        L.orc_fingerprint_blocks_mt.argtypes = [u8p, sz, sz, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int,
                                                C.c_void_p, C.c_void_p]
        L.orc_bench_decompress_mt.restype = C.c_double
        L.orc_bench_decompress_mt.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), sz,
                                              C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, err):
        self.err = err.tuple() if isinstance(err, OrcError) else err
        super().__init__(repr(self.err))


def max_compress_len(n):
    return lib().orc_max_compress_len(n)


def compress(data: bytes) -> bytes:
    data = bytes(data)
    cap = max_compress_len(len(data))
    buf = C.create_string_buffer(max(cap, 1))
    n, e = C.c_size_t(0), OrcError()
    rc = lib().orc_compress(data, len(data), buf, cap, C.byref(n), C.byref(e))
    if rc:
        raise OracleError(e)
    return buf.raw[:n.value]


def decompress_len(data: bytes) -> int:
    n, e = C.c_size_t(0), OrcError()
    rc = lib().orc_decompress_len(bytes(data), len(data), C.byref(n), C.byref(e))
    if rc:
        raise OracleError(e)
    return n.value


def decompress(data: bytes, cap=None) -> bytes:
    data = bytes(data)
    if cap is None:
        cap = decompress_len(data)
    buf = C.create_string_buffer(max(cap, 1))
    n, e = C.c_size_t(0), OrcError()
    rc = lib().orc_decompress(data, len(data), buf, cap, C.byref(n), C.byref(e))
    if rc:
        raise OracleError(e)
    return buf.raw[:n.value]


def crc32c(data: bytes) -> int:
    return lib().orc_crc32c(bytes(data), len(data))


def crc32c_masked(data: bytes) -> int:
    return lib().orc_crc32c_masked(bytes(data), len(data))


def compress_frame(data: bytes) -> bytes:
    buf = C.create_string_buffer(8 + 76490)
    n = C.c_size_t(0)
    rc = lib().orc_compress_frame(bytes(data), len(data), buf, C.byref(n))
    if rc:
        raise OracleError(("Assert", 0, 0, 0))
    return buf.raw[:n.value]


def frame_encode(data: bytes) -> bytes:
    data = bytes(data)
    cap = lib().orc_frame_max_len(len(data))
    buf = C.create_string_buffer(max(cap, 1))
    n = C.c_size_t(0)
    rc = lib().orc_frame_encode(data, len(data), buf, cap, C.byref(n))
    if rc:
        raise OracleError(("Assert", 0, 0, 0))
    return buf.raw[:n.value]


def frame_decode(data: bytes) -> bytes:
    data = bytes(data)
    n, e = C.c_size_t(0), OrcError()
    lib().orc_frame_decode(data, len(data), None, 0, C.byref(n), C.byref(e))
    cap = n.value
    buf = C.create_string_buffer(max(cap, 1))
    rc = lib().orc_frame_decode(data, len(data), buf, cap, C.byref(n), C.byref(e))
    if rc:
        raise OracleError(e)
    return buf.raw[:n.value]


def bench_compress_mt(text: bytes, block_len: int, first: int, count: int, mul: int, threads: int):
    tot = C.c_uint64(0)
    secs = lib().orc_bench_compress_mt(text, len(text), block_len, first, count, mul, threads, C.byref(tot))
    return secs, tot.value


def fingerprint_blocks_mt(text: bytes, block_len: int, base: int, step: int, count: int, mul: int, threads: int):
    """(seconds, lens, crcs) of generator blocks base, base+step, ...: compressed length and masked CRC-32C of the
    compressed stream (bench.py's full-coverage parity check)."""
    import numpy as np
    lens = np.zeros(max(count, 1), dtype=np.uint32)
    crcs = np.zeros(max(count, 1), dtype=np.uint32)
    secs = lib().orc_fingerprint_blocks_mt(text, len(text), block_len, base, step, count, mul, threads,
                                           lens.ctypes.data, crcs.ctypes.data)
    return secs, lens[:count], crcs[:count]


def bench_decompress_mt(streams, count: int, threads: int):
    arr = (C.c_char_p * len(streams))(*streams)
    lens = (C.c_size_t * len(streams))(*[len(s) for s in streams])
    tot = C.c_uint64(0)
    secs = lib().orc_bench_decompress_mt(arr, lens, len(streams), count, threads, C.byref(tot))
    return secs, tot.value
