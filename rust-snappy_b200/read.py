"""`snap::read::{FrameDecoder, FrameEncoder}` mirrored (reference src/read.rs).

FrameDecoder keeps the reference's chunk state machine (src/read.rs:104-239):
one chunk is decoded per refill; the raw decode + checksum run on the GPU.
FrameEncoder (src/read.rs:272-410) turns each underlying read() of <=64KB into
one chunk, so its output equals write::FrameEncoder's for slice readers.

`FrameDecoder(rdr, batch_chunks=N)` (default 1 = the reference's behaviour: exactly
one chunk is pulled from the reader per refill) reads AHEAD up to N data chunks and
decodes them with one device call (header walk, K2, checksum); bytes and errors come
out in the same order, but the underlying reader is consumed earlier.
"""
from . import frame
from .error import Error, UnexpectedEof
from .raw import Decoder, crc32c_masked, decompress_len


def _read_exact(r, n):
    out = bytearray()
    while len(out) < n:
        piece = r.read(n - len(out))
        if not piece:
            raise UnexpectedEof("failed to fill whole buffer")
        out += piece
    return bytes(out)


class FrameDecoder:
    def __init__(self, rdr, batch_chunks=1):
        self._r = rdr
        self._batch = max(1, int(batch_chunks))
        self._pending = None             # error raised once the bytes decoded before it were served (batch mode)
        self._eof = False
        self._dec = Decoder()
        self._src = bytearray(frame.MAX_COMPRESS_BLOCK_SIZE)
        self._dst = b""
        self._dsts = 0
        self._read_stream_ident = False

    def get_ref(self):
        return self._r

    get_mut = get_ref
    into_inner = get_ref

    def read(self, size=-1) -> bytes:
        if size is None or size < 0:
            parts = []
            while True:
                p = self.read(1 << 20)
                if not p:
                    return b"".join(parts)
                parts.append(p)
        if size == 0:
            return b""
        if self._batch > 1:
            return self._read_batched(size)
        while True:
            if self._dsts < len(self._dst):
                out = self._dst[self._dsts:self._dsts + size]
                self._dsts += len(out)
                return out
            head = self._r.read(4)
            if not head:
                return b""                                   # clean EOF (src/read.rs:119-121)
            if len(head) < 4:
                head += _read_exact(self._r, 4 - len(head))
            self._src[0:4] = head
            ty = head[0]
            if not self._read_stream_ident:
                if ty != 0xFF:
                    raise Error("StreamHeader", byte=ty)
                self._read_stream_ident = True
            ln = head[1] | (head[2] << 8) | (head[3] << 16)
            if ln > len(self._src):
                raise Error("UnsupportedChunkLength", len=ln, header=False)
            if 0x02 <= ty <= 0x7F:
                raise Error("UnsupportedChunkType", byte=ty)
            if 0x80 <= ty <= 0xFE:                           # reserved-skippable and padding
                self._src[0:ln] = _read_exact(self._r, ln)
            elif ty == 0xFF:
                if ln != len(frame.STREAM_BODY):
                    raise Error("UnsupportedChunkLength", len=ln, header=True)
                body = _read_exact(self._r, ln)
                self._src[0:ln] = body
                if body != frame.STREAM_BODY:
                    raise Error("StreamHeaderMismatch", bytes=body)
            else:
                if ln < 4:
                    raise Error("UnsupportedChunkLength", len=ln, header=False)
                expected = int.from_bytes(_read_exact(self._r, 4), "little")
                n = ln - 4
                if ty == 0x01:
                    if n > frame.MAX_BLOCK_SIZE:
                        raise Error("UnsupportedChunkLength", len=n, header=False)
                    data = _read_exact(self._r, n)
                else:
                    self._src[0:n] = _read_exact(self._r, n)
                    dn = decompress_len(self._src)           # whole buffer, like src/read.rs:216
                    if dn > frame.MAX_BLOCK_SIZE:
                        raise Error("UnsupportedChunkLength", len=dn, header=False)
                    out = bytearray(dn)
                    self._dec.decompress(bytes(self._src[0:n]), out)
                    data = bytes(out)
                got = crc32c_masked(data)
                if expected != got:
                    raise Error("Checksum", expected=expected, got=got)
                self._dst = data
                self._dsts = 0

    def read_to_end(self) -> bytes:
        return self.read(-1)

    # -- batch mode: raw chunks are collected (each read from the underlying reader is exactly the one the reference
    # would issue, only sooner) and handed to the device decoder as one stream; a stream identifier in front of a later
    # batch stands in for the one this decoder has already seen (identifier chunks may repeat, src/read.rs:166-178)
    def _read_batched(self, size) -> bytes:
        while True:
            if self._dsts < len(self._dst):
                out = self._dst[self._dsts:self._dsts + size]
                self._dsts += len(out)
                return out
            if self._pending is not None:
                e, self._pending = self._pending, None
                raise e
            if self._eof:
                return b""
            seen = self._read_stream_ident
            raw = bytearray(frame.STREAM_IDENTIFIER) if seen else bytearray()
            base = len(raw)
            chunks = 0
            while chunks < self._batch and not self._eof:
                head = _read_upto(self._r, 4)
                raw += head
                if len(head) < 4:
                    self._eof = True                         # clean end (0 bytes) or a truncated header
                    break
                self._read_stream_ident = True
                ln = head[1] | (head[2] << 8) | (head[3] << 16)
                if ln > frame.MAX_COMPRESS_BLOCK_SIZE or 0x02 <= head[0] <= 0x7F:
                    self._eof = True                         # the decoder stops here with the reference's error
                    break
                body = _read_upto(self._r, ln)
                raw += body
                if len(body) < ln:
                    self._eof = True
                    break
                if head[0] in (0x00, 0x01):
                    chunks += 1
            self._dst, self._dsts = b"", 0
            if len(raw) > base:
                self._dst, self._pending = frame.decode_all_partial(bytes(raw))


def _read_upto(r, n):
    """Like read_exact, but a short stream returns what there was."""
    out = bytearray()
    while len(out) < n:
        piece = r.read(n - len(out))
        if not piece:
            break
        out += piece
    return bytes(out)


class FrameEncoder:
    _MAX_BLOCK = 10 + 8 + frame.MAX_COMPRESS_BLOCK_SIZE     # src/read.rs:33-35

    def __init__(self, rdr):
        self._r = rdr
        self._dst = b""
        self._dsts = 0
        self._wrote_stream_ident = False

    def get_ref(self):
        return self._r

    get_mut = get_ref

    def _read_frame(self) -> bytes:
        src = self._r.read(frame.MAX_BLOCK_SIZE)             # ONE underlying read (src/read.rs:378-381)
        if not src:
            return b""
        ident = not self._wrote_stream_ident
        self._wrote_stream_ident = True
        return frame.encode_chunks(src, include_ident=ident)

    def read(self, size=-1) -> bytes:
        if size is None or size < 0:
            parts = []
            while True:
                p = self.read(1 << 20)
                if not p:
                    return b"".join(parts)
                parts.append(p)
        if size == 0:
            return b""
        if self._dsts >= len(self._dst):
            self._dst = self._read_frame()
            self._dsts = 0
        out = self._dst[self._dsts:self._dsts + size]
        self._dsts += len(out)
        return out

    def read_to_end(self) -> bytes:
        return self.read(-1)
