"""Multi-GPU sharding of the frame path (BASELINE configs[4]).

Frame chunks are self-contained (reference src/frame.rs:62-104), so a stream of
`nchunks` 64KB chunks is cut into contiguous chunk ranges, one per rank; every
rank encodes its range with the CUDA kernels (`sb_frame_encode_device`, stream
identifier on rank 0 only). The single exchange step reassembles the framed
output: an all-gather of the per-rank compressed byte counts (-> every rank's
global offset), then an all-gather of the payload, padded to the largest rank.
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


def all_gather_stream(local, dist, group=None):
    """Reassemble the framed stream on every rank.

    `local` is a 1-D uint8 tensor holding this rank's part (cpu tensor with gloo,
    cuda tensor with nccl). Returns (full_stream_tensor, offsets, sizes).
    """
    import torch
    world = dist.get_world_size(group)
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    offs, total = global_offsets(sizes)
    pad = max(sizes) if sizes else 0
    mine = torch.zeros(pad, dtype=torch.uint8, device=local.device)
    mine[:local.numel()] = local
    parts = [torch.empty(pad, dtype=torch.uint8, device=local.device) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    full = torch.empty(total, dtype=torch.uint8, device=local.device)
    for r in range(world):
        full[offs[r]:offs[r] + sizes[r]] = parts[r][:sizes[r]]
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
