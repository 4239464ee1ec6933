"""`snap::write::FrameEncoder` mirrored (reference src/write.rs:34-192).

Same buffering rules as the reference -- they decide where chunk boundaries
fall, so they are part of the byte-exact contract: a 64KB staging buffer `src`;
a write larger than the free space goes straight to `Inner.write` when `src` is
empty (src/write.rs:132-135), otherwise it first tops `src` up and flushes it.
"""
from . import frame


class FrameEncoder:
    def __init__(self, wtr):
        self._w = wtr
        self._src = bytearray()
        self._wrote_stream_ident = False
        self._inner_taken = False

    # -- Inner::write (src/write.rs:165-192): stream identifier once, then chunks
    def _inner_write(self, buf) -> int:
        ident = not self._wrote_stream_ident
        self._wrote_stream_ident = True
        if ident:
            self._w.write(frame.STREAM_IDENTIFIER)
        if len(buf):
            self._w.write(frame.encode_chunks(buf, include_ident=False))
        return len(buf)

    def write(self, buf) -> int:
        buf = memoryview(buf).cast("B")
        total = 0
        while True:
            free = frame.MAX_BLOCK_SIZE - len(self._src)
            if len(buf) <= free:
                break
            if not self._src:
                n = self._inner_write(buf)
            else:
                self._src += buf[:free]
                self.flush()
                n = free
            buf = buf[n:]
            total += n
        self._src += buf
        return total + len(buf)

    def write_all(self, buf):
        self.write(buf)

    def flush(self):
        if not self._src:
            return
        self._inner_write(bytes(self._src))
        del self._src[:]

    def into_inner(self):
        self.flush()
        self._inner_taken = True
        return self._w

    def get_ref(self):
        return self._w

    get_mut = get_ref

    def close(self):
        if not self._inner_taken:
            self.flush()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()       # Drop flushes, ignoring errors (src/write.rs:112-120)
        return False
