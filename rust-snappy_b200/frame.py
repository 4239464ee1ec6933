"""Frame-format constants and one-shot helpers (reference src/frame.rs)."""
import ctypes as C

from . import _lib
from .error import from_c
from .raw import _ptr

MAX_BLOCK_SIZE = 1 << 16                      # src/lib.rs:97
MAX_COMPRESS_BLOCK_SIZE = 76490               # src/frame.rs:12
STREAM_IDENTIFIER = b"\xFF\x06\x00\x00sNaPpY"  # src/frame.rs:18
STREAM_BODY = b"sNaPpY"
CHUNK_HEADER_AND_CRC_SIZE = 8                 # src/frame.rs:26


def encode_chunks(data, include_ident: bool) -> bytes:
    """Chunks for `data` exactly as write::Inner::write emits them (src/write.rs:165-192)."""
    n = len(data)
    if n == 0:
        return STREAM_IDENTIFIER if include_ident else b""
    cap = _lib.lib().sb_frame_max_len(n)
    out = bytearray(cap)
    ip, k1 = _ptr(data)
    op, k2 = _ptr(out)
    m, e = C.c_size_t(0), _lib.SbError()
    if _lib.lib().sb_frame_encode_ex(ip, n, op, cap, C.byref(m), 1 if include_ident else 0, C.byref(e)):
        raise from_c(e)
    return bytes(out[:m.value])


def decode_all(stream) -> bytes:
    """read::FrameDecoder::new(stream).read_to_end() (src/read.rs:104-239)."""
    n = len(stream)
    ip, k1 = _ptr(stream) if n else (None, None)
    m, e = C.c_size_t(0), _lib.SbError()
    L = _lib.lib()
    if L.sb_frame_decode(ip, n, None, 0, C.byref(m), C.byref(e)):
        raise from_c(e)
    out = bytearray(max(m.value, 1))
    op, k2 = _ptr(out)
    if L.sb_frame_decode(ip, n, op, m.value, C.byref(m), C.byref(e)):
        raise from_c(e)
    return bytes(out[:m.value])


def decode_all_partial(stream):
    """Like decode_all, but a failing stream yields (bytes produced before the failure, the exception) instead of
    raising: what a reader delivers before its n-th read fails."""
    n = len(stream)
    ip, k1 = _ptr(stream) if n else (None, None)
    m, e = C.c_size_t(0), _lib.SbError()
    L = _lib.lib()
    if L.sb_frame_decode(ip, n, None, 0, C.byref(m), C.byref(e)):
        return b"", from_c(e)
    out = bytearray(max(m.value, 1))
    op, k2 = _ptr(out)
    rc = L.sb_frame_decode(ip, n, op, m.value, C.byref(m), C.byref(e))
    return bytes(out[:m.value]), (from_c(e) if rc else None)
