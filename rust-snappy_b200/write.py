"""`snap::write::FrameEncoder` mirrored (reference src/write.rs:34-192).

Same buffering rules as the reference -- they decide where chunk boundaries
fall, so they are part of the byte-exact contract: a 64KB staging buffer `src`;
a write larger than the free space goes straight to `Inner.write` when `src` is
empty (src/write.rs:132-135), otherwise it first tops `src` up and flushes it.

`batch_chunks=N` (default 1 = the reference's behaviour: every chunk reaches the
writer as soon as it is complete) queues up to N full 64KB chunks and encodes them
with ONE device call; a partial chunk, flush(), into_inner() and close() drain the
queue. The bytes written are identical -- only the moment they reach the writer moves.
"""
from . import frame


class FrameEncoder:
    def __init__(self, wtr, batch_chunks=1):
        self._w = wtr
        self._src = bytearray()
        self._wrote_stream_ident = False
        self._inner_taken = False
        self._batch = max(1, int(batch_chunks))
        self._queue = bytearray()        # full chunks waiting for the next device call (batch mode)

    def _drain(self):
        if self._queue:
            self._w.write(frame.encode_chunks(self._queue, include_ident=False))
            del self._queue[:]

    # -- Inner::write (src/write.rs:165-192): stream identifier once, then chunks
    def _inner_write(self, buf) -> int:
        ident = not self._wrote_stream_ident
        self._wrote_stream_ident = True
        if ident:
            self._w.write(frame.STREAM_IDENTIFIER)
        if len(buf):
            if self._batch == 1:
                self._w.write(frame.encode_chunks(buf, include_ident=False))
            else:
                # chunk boundaries inside `buf` are every 64KB with the partial chunk last: queued full chunks followed
                # by `buf` encode to the same bytes in one call as chunk by chunk
                self._queue += buf
                if len(buf) % frame.MAX_BLOCK_SIZE or len(self._queue) >= self._batch * frame.MAX_BLOCK_SIZE:
                    self._drain()
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
        if self._src:
            self._inner_write(bytes(self._src))
            del self._src[:]
        self._drain()

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
