"""Side measurement (not the headline bench): frame-format encode of a device-resident synthetic
stream (BASELINE configs[3] shape, scaled to --gib) + the CRC-32C kernel alone. Prints one JSON line."""
import argparse, ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as graft
from bench import load_text, BLOCK, MUL

ap = argparse.ArgumentParser()
ap.add_argument("--gib", type=float, default=4.0)
ap.add_argument("--verify", action="store_true")
args = ap.parse_args()
snap = graft.load_package()
L = snap._lib.lib()
err = snap._lib.SbError()
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
text = load_text()
nchunks = int(args.gib * (1 << 30)) // BLOCK
n = nchunks * BLOCK
t_text = torch.frombuffer(bytearray(text), dtype=torch.uint8).to(dev)
t_in = torch.empty(n, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
assert L.sb_generate_blocks_device(t_text.data_ptr(), len(text), t_in.data_ptr(), BLOCK, BLOCK, 0, nchunks, MUL, st, C.byref(err)) == 0
cap = L.sb_frame_max_len(n)
t_out = torch.empty(cap, dtype=torch.uint8, device=dev)
total = C.c_uint64(0)
def enc():
    rc = L.sb_frame_encode_device(t_in.data_ptr(), n, t_out.data_ptr(), cap, 1, C.byref(total), st, C.byref(err))
    assert rc == 0, err.code
enc(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    enc()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
# CRC kernel alone
crc = torch.zeros(nchunks, dtype=torch.int32, device=dev)
b = snap._lib.SbBatch()
b.in_base, b.in_stride, b.in_len_uniform, b.out_lens, b.count = t_in.data_ptr(), BLOCK, BLOCK, crc.data_ptr(), nchunks
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
L.sb_crc32c_masked_batch_device(C.byref(b), st, C.byref(err)); torch.cuda.synchronize()
e0.record()
for _ in range(3):
    L.sb_crc32c_masked_batch_device(C.byref(b), st, C.byref(err))
e1.record(); torch.cuda.synchronize()
crc_ms = e0.elapsed_time(e1) / 3
res = {"frame_encode_device": {"bytes": n, "chunks": nchunks, "stream_bytes": total.value, "seconds": dt, "uncompressed_gbs": n / dt / 1e9},
       "crc32c_kernel": {"ms": crc_ms, "gbs": n / (crc_ms / 1e3) / 1e9}}
if args.verify:
    from oracle import oracle as orc
    k = min(nchunks, 64)
    host = bytes(t_in[:k * BLOCK].cpu().numpy())
    want = orc.frame_encode(host)
    got = bytes(t_out[:len(want)].cpu().numpy())
    res["verify_first_chunks_equal_oracle"] = got == want
    full = bytes(t_out[:total.value].cpu().numpy()) if total.value < (1 << 31) else None
    if full is not None:
        res["decode_roundtrip_ok"] = snap.frame.decode_all(full) == bytes(t_in.cpu().numpy())
print(json.dumps(res))
