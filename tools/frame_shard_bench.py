"""BASELINE configs[4] shape at reduced scale: a framed stream of 64KB chunks is split across ranks
(contiguous chunk ranges), every rank frame-encodes its range on its GPU (stream identifier on rank 0
only), per-rank sizes and the payload are all-gathered over NCCL and the reassembled stream is checked:
  * --verify-mib M : the first M MiB-stream is compared byte for byte with the oracle's single-stream
    encoding (small case), 
  * the full case reports encode-only and encode+all-gather throughput (device timed, max over ranks).
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/frame_shard_bench.py
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import torch
import torch.distributed as dist
import __graft_entry__ as graft
from bench import load_text, BLOCK, MUL

ap = argparse.ArgumentParser()
ap.add_argument("--gib-per-rank", type=float, default=2.0)
ap.add_argument("--verify-mib", type=int, default=48)
args = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
snap = graft.load_package()
L = snap._lib.lib()
err = snap._lib.SbError()
text = load_text()
t_text = torch.frombuffer(bytearray(text), dtype=torch.uint8).to(dev)
st = torch.cuda.current_stream().cuda_stream


def gen(first_chunk, nchunks):
    t = torch.empty(nchunks * BLOCK, dtype=torch.uint8, device=dev)
    assert L.sb_generate_blocks_device(t_text.data_ptr(), len(text), t.data_ptr(), BLOCK, BLOCK, first_chunk, nchunks, MUL, st, C.byref(err)) == 0
    return t


def run(total_chunks, verify):
    lo, hi = snap.shard.chunk_range(total_chunks, rank, world)
    mine = gen(lo, hi - lo)                              # this rank's chunk range of the global stream
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    local_part = snap.shard.encode_device(mine, include_ident=(rank == 0))
    e1.record()
    if world > 1:
        full, offs, sizes = snap.shard.all_gather_stream(local_part, dist)
    else:
        full, offs, sizes = local_part, [0], [local_part.numel()]
    e2.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1), e0.elapsed_time(e2)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = None
    if verify and rank == 0:
        from oracle import oracle as orc
        span = len(text) - BLOCK
        data = b"".join(text[(i * MUL) % span:][:BLOCK] for i in range(total_chunks))
        ok = bytes(full.cpu().numpy()) == orc.frame_encode(data)
    return t.tolist(), int(full.numel()), ok


ms_small, n_small, ok = run(args.verify_mib * 16, True)
chunks = int(args.gib_per_rank * (1 << 30)) // BLOCK * world
run(chunks, False)                                       # warm-up (allocations)
ms, stream_bytes, _ = run(chunks, False)
if rank == 0:
    u = chunks * BLOCK
    print(json.dumps({"n_gpus": world, "chunks": chunks, "uncompressed_bytes": u, "stream_bytes": stream_bytes,
                      "verify_small_stream_equals_oracle": ok,
                      "encode_only_gbs": u / (ms[0] / 1e3) / 1e9, "encode_plus_allgather_gbs": u / (ms[1] / 1e3) / 1e9,
                      "ms_encode": ms[0], "ms_total": ms[1]}))
if world > 1:
    dist.destroy_process_group()
