"""N>1 host logic on CPU: world_size-2 gloo. The per-rank encoder is the oracle here
(no GPU in this container); the GPU path swaps in sb_frame_encode_device and is
covered by test_gpu_sharded_frame_encode (needs 1 GPU) and bench.py --gpus N."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT, corpus


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, data, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import __graft_entry__ as g
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    snap = g.load_package()
    t = torch.frombuffer(bytearray(data), dtype=torch.uint8)

    def enc(lo, hi, ident):
        body = orc.frame_encode(data[lo:hi]) if hi > lo else b""
        if not ident:
            body = body[10:]
        return torch.frombuffer(bytearray(body), dtype=torch.uint8) if body else torch.empty(0, dtype=torch.uint8)

    full, offs, sizes = snap.shard.frame_encode_sharded(t, rank, world, dist=dist, encode_range=enc)
    q.put((rank, bytes(full.numpy()), offs, sizes))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,cut", [("html_x_4", None), ("alice29.txt", 70000), ("urls.10K", 65536)])
def test_sharded_frame_stream_equals_single_stream(oracle, name, cut):
    data = corpus(name)[:cut] if cut else corpus(name)
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, data, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    want = oracle.frame_encode(data)
    for rank, full, offs, sizes in res:
        assert full == want, "rank %d reassembled a different stream" % rank
        assert sum(sizes) == len(want) and offs[0] == 0 and offs[1] == sizes[0]


def test_chunk_ranges_cover_exactly():
    import __graft_entry__ as g
    shard = g.load_package().shard
    for n in [0, 1, 2, 7, 8, 9, 16777216, 1048577]:
        for world in [1, 2, 4, 8]:
            ranges = [shard.chunk_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert max(h - l for l, h in ranges) - min(h - l for l, h in ranges) <= 1
