"""`snap::raw` mirrored over the C ABI (reference src/raw.rs:13-14).

`Encoder.compress/compress_vec`, `Decoder.decompress/decompress_vec`,
`max_compress_len`, `decompress_len` keep the reference's argument meaning and
error behaviour (src/compress.rs:42-169, src/decompress.rs:30-110); the work is
done by the CUDA kernels behind `sb_compress` / `sb_decompress`.
"""
import ctypes as C

from . import _lib
from .error import from_c


def _ptr(buf):
    """address of a bytes-like object without copying (bytes, bytearray, memoryview, numpy)."""
    if isinstance(buf, bytes):
        return C.cast(C.c_char_p(buf), C.c_void_p).value, buf
    mv = memoryview(buf)
    if mv.readonly:
        b = bytes(mv)
        return C.cast(C.c_char_p(b), C.c_void_p).value, b
    arr = (C.c_char * mv.nbytes).from_buffer(mv)
    return C.addressof(arr), arr


def max_compress_len(input_len: int) -> int:
    return _lib.lib().sb_max_compress_len(input_len)


def decompress_len(data) -> int:
    p, keep = _ptr(data)
    n, e = C.c_size_t(0), _lib.SbError()
    if _lib.lib().sb_decompress_len(p, len(data), C.byref(n), C.byref(e)):
        raise from_c(e)
    return n.value


class Encoder:
    """snap::raw::Encoder (src/compress.rs:67-170)."""

    def compress(self, input, output) -> int:
        ip, k1 = _ptr(input)
        op, k2 = _ptr(output)
        n, e = C.c_size_t(0), _lib.SbError()
        if _lib.lib().sb_compress(ip, len(input), op, len(output), C.byref(n), C.byref(e)):
            raise from_c(e)
        return n.value

    def compress_vec(self, input) -> bytes:
        buf = bytearray(max(max_compress_len(len(input)), 1))
        n = self.compress(input, buf)
        return bytes(buf[:n])


class Decoder:
    """snap::raw::Decoder (src/decompress.rs:45-111)."""

    def decompress(self, input, output) -> int:
        ip, k1 = _ptr(input)
        op, k2 = _ptr(output) if len(output) else (None, None)
        n, e = C.c_size_t(0), _lib.SbError()
        if _lib.lib().sb_decompress(ip, len(input), op, len(output), C.byref(n), C.byref(e)):
            raise from_c(e)
        return n.value

    def decompress_vec(self, input) -> bytes:
        buf = bytearray(decompress_len(input))
        n = self.decompress(input, buf)
        return bytes(buf[:n])


def crc32c_masked(data) -> int:
    """crc32::CheckSummer::crc32c_masked (src/crc32.rs:35-38), computed on the GPU."""
    p, keep = _ptr(data) if len(data) else (None, None)
    out, e = C.c_uint32(0), _lib.SbError()
    if _lib.lib().sb_crc32c_masked(p, len(data), C.byref(out), C.byref(e)):
        raise from_c(e)
    return out.value
