"""Multi-GPU sharding of the frame path (BASELINE configs[4]).

Frame chunks are self-contained (reference src/frame.rs:62-104), so a stream of
64KB chunks is cut into contiguous chunk ranges, one per rank; every rank encodes
its range with the CUDA kernels (stream identifier on rank 0 only). The single
exchange step reassembles the framed output:

  1. all-gather of the per-rank compressed byte counts (one int64 per rank, taken
     straight from the device-side result record) -> every rank's offset;
  2. all-gather-v of the payload: grouped send/recv (`batch_isend_irecv`) directly
     into the pre-sized destination at those offsets -- no padding to the largest
     rank, no staging copies, no Python reassembly loop.

Large streams run in waves (`WavePipeline`): wave w covers the global chunk range
[w*world*W, (w+1)*world*W) and rank r takes its r-th slice, so the stream order
inside a wave is the rank order and a wave's offset is known as soon as the
previous waves' sizes are. The payload exchange of wave k is issued asynchronously
and overlaps the kernels of wave k+1.

One process per GPU; `torch.distributed` (NCCL on GPUs, gloo in the CPU tests)
is plumbing only.
"""
import ctypes as C


def chunk_range(nchunks: int, rank: int, world: int):
    """Contiguous, balanced chunk range [lo, hi) of `rank` (first ranks take the remainder)."""
    base, rem = divmod(nchunks, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_offsets(sizes):
    """Exclusive scan of per-rank compressed sizes -> (offset of every rank, total)."""
    offs, at = [], 0
    for s in sizes:
        offs.append(at)
        at += int(s)
    return offs, at


def gather_sizes(n_local, dist, device, group=None):
    """All-gather one int64 per rank. `n_local` is an int or a 1-element int64 tensor on `device`."""
    import torch
    world = dist.get_world_size(group)
    mine = n_local if torch.is_tensor(n_local) else torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    allv = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allv, mine.reshape(1), group=group)
    return [int(x) for x in allv.cpu().tolist()]


def exchange_payload(local, sizes, out, base, dist, group=None, async_op=False):
    """All-gather-v: every rank's `local[:sizes[rank]]` lands at out[base + offs[r] : ...] on every rank.

    Grouped point-to-point operations straight into the destination; the caller's own part is a local copy.
    Returns the list of outstanding work handles when async_op is set (call .wait() on each before reusing
    `local` or reading `out`)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    offs, _total = global_offsets(sizes)
    mine = local[:sizes[rank]]
    out[base + offs[rank]: base + offs[rank] + sizes[rank]].copy_(mine, non_blocking=True)
    ops = []
    for k in range(1, world):
        dst, src = (rank + k) % world, (rank - k) % world
        if sizes[rank]:
            ops.append(dist.P2POp(dist.isend, mine, dst, group))
        if sizes[src]:
            ops.append(dist.P2POp(dist.irecv, out[base + offs[src]: base + offs[src] + sizes[src]], src, group))
    works = dist.batch_isend_irecv(ops) if ops else []
    if async_op:
        return works
    for w in works:
        w.wait()
    return []


def all_gather_stream(local, dist, group=None):
    """Reassemble the framed stream on every rank.

    `local` is a 1-D uint8 tensor holding this rank's part (cpu tensor with gloo,
    cuda tensor with nccl). Returns (full_stream_tensor, offsets, sizes).
    """
    import torch
    sizes = gather_sizes(local.numel(), dist, local.device, group)
    offs, total = global_offsets(sizes)
    full = torch.empty(total, dtype=torch.uint8, device=local.device)
    exchange_payload(local, sizes, full, 0, dist, group)
    return full, offs, sizes


def frame_encode_sharded(data, rank, world, dist=None, group=None, encode_range=None):
    """Encode this rank's chunk range of `data` (1-D uint8 tensor, whole stream visible or
    synthesised per rank) and, when `dist` is given, all-gather the framed stream.

    `encode_range(lo_byte, hi_byte, include_ident) -> 1-D uint8 tensor` defaults to the
    CUDA path (device tensor in, device tensor out).
    """
    n = data.numel()
    nchunks = (n + 65535) // 65536
    lo, hi = chunk_range(nchunks, rank, world)
    lo_b, hi_b = lo * 65536, min(hi * 65536, n)
    if encode_range is None:
        encode_range = lambda a, b, ident: encode_device(data[a:b], ident)   # noqa: E731
    local = encode_range(lo_b, hi_b, rank == 0)
    if dist is None or world == 1:
        return local, [0], [local.numel()]
    return all_gather_stream(local, dist, group)


def encode_device(t, include_ident):
    """sb_frame_encode_device over a CUDA uint8 tensor; returns a CUDA uint8 tensor."""
    import torch
    from . import _lib
    from .error import from_c
    L = _lib.lib()
    n = t.numel()
    cap = L.sb_frame_max_len(n)
    out = torch.empty(max(cap, 1), dtype=torch.uint8, device=t.device)
    total, err = C.c_uint64(0), _lib.SbError()
    if n == 0:
        return out[:0]
    rc = L.sb_frame_encode_device(t.data_ptr(), n, out.data_ptr(), cap, 1 if include_ident else 0, C.byref(total),
                                  torch.cuda.current_stream().cuda_stream, C.byref(err))
    if rc:
        raise from_c(err)
    return out[:total.value]


class WavePipeline:
    """Wave-by-wave sharded frame encode with the exchange of wave k overlapping the kernels of wave k+1.

    Every rank calls encode(w, d_in) for w = 0, 1, ...: the wave's chunks are frame-encoded stream-ordered
    (sb_frame_encode_device_ws, no host synchronisation), then the PREVIOUS wave is finished: its size is
    all-gathered from the device-side result record and its payload exchange is issued asynchronously into the
    gather buffer of that wave (two buffers alternate). flush() finishes the last wave. `on_wave(w, buf, offs,
    sizes, total)` (optional) sees each reassembled wave after its exchange completed.
    """

    def __init__(self, wave_bytes, dist, device, exchange=True, on_wave=None):
        import torch
        from . import _lib
        self.L, self._lib, self.torch = _lib.lib(), _lib, torch
        self.dist, self.dev, self.exchange, self.on_wave = dist, device, exchange, on_wave
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.cap = self.L.sb_frame_max_len(wave_bytes)
        self.scratch_bytes = self.L.sb_frame_encode_scratch_bytes(wave_bytes)
        self.scratch = torch.empty(self.scratch_bytes + 256, dtype=torch.uint8, device=device)
        self.local = [torch.empty(self.cap + 16, dtype=torch.uint8, device=device) for _ in range(2)]
        self.result = [torch.zeros(8, dtype=torch.int64, device=device) for _ in range(2)]   # sb_frame_result (48 bytes)
        gcap = self.cap * self.world if exchange else 0
        self.gathered = [torch.empty(gcap + 16, dtype=torch.uint8, device=device) for _ in range(2)] if exchange and self.world > 1 else None
        self.err = _lib.SbError()
        if self.gathered is not None:
            # bring the point-to-point transports up now (NCCL connects peers lazily at the first send/recv): one byte
            # to and from every peer, so that no connection setup runs beside the first waves' kernels
            warm = torch.zeros(self.world, dtype=torch.uint8, device=device)
            exchange_payload(warm, [1] * self.world, self.gathered[0], 0, dist)
            torch.cuda.synchronize()
        self.pending = None          # (wave, buffer index) encoded but not exchanged yet
        self.works = [[], []]        # outstanding exchanges per buffer
        self.stream_bytes = 0        # bytes of the reassembled stream so far (all ranks, all finished waves)
        self.nccl_bytes = 0          # bytes this rank received over the fabric

    def encode(self, w, d_in, nbytes):
        b = w & 1
        for wk in self.works[b]:      # wave w-2 used these buffers: its exchange must be done
            wk.wait()
        self.works[b] = []
        st = self.torch.cuda.current_stream().cuda_stream
        rc = self.L.sb_frame_encode_device_ws(d_in, nbytes, self.local[b].data_ptr(), self.cap, 1 if (w == 0 and self.rank == 0) else 0,
                                              None, self.result[b].data_ptr(), self.scratch.data_ptr(), self.scratch_bytes + 256, st,
                                              C.byref(self.err))
        if rc:
            from .error import from_c
            raise from_c(self.err)
        prev, self.pending = self.pending, (w, b)
        if prev is not None:
            self._finish(*prev)

    def _finish(self, w, b):
        if self.world == 1 or not self.exchange:
            if self.world == 1 and self.on_wave is not None:
                n = int(self.result[b][4].item())
                self.on_wave(w, self.local[b], [0], [n], n)
            return
        sizes = gather_sizes(self.result[b][4:5], self.dist, self.dev)     # result.bytes
        offs, total = global_offsets(sizes)
        self.works[b] = exchange_payload(self.local[b], sizes, self.gathered[b], 0, self.dist, async_op=True)
        self.stream_bytes += total
        self.nccl_bytes += total - sizes[self.rank]
        if self.on_wave is not None:
            for wk in self.works[b]:
                wk.wait()
            self.works[b] = []
            self.on_wave(w, self.gathered[b], offs, sizes, total)

    def flush(self):
        if self.pending is not None:
            self._finish(*self.pending)
            self.pending = None
        for b in range(2):
            for wk in self.works[b]:
                wk.wait()
            self.works[b] = []
