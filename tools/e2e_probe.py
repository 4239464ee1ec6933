"""Times the two host-batch C ABI calls separately + raw pinned PCIe copies (diagnostic)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as graft
from bench import load_text, BLOCK, MUL
snap = graft.load_package(); L = snap._lib.lib(); err = snap._lib.SbError()
torch.cuda.set_device(0); dev = torch.device("cuda:0")
n = 131072
text = load_text()
t_text = torch.frombuffer(bytearray(text), dtype=torch.uint8).to(dev)
t_in = torch.empty(n * BLOCK, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
L.sb_generate_blocks_device(t_text.data_ptr(), len(text), t_in.data_ptr(), BLOCK, BLOCK, 0, n, MUL, st, C.byref(err))
h_in = torch.empty(n * BLOCK, dtype=torch.uint8).pin_memory(); h_in.copy_(t_in)
h_out = torch.empty(n * BLOCK, dtype=torch.uint8).pin_memory()
cap = int(L.sb_max_compress_len(BLOCK))
h_c = torch.empty(n * 45000, dtype=torch.uint8).pin_memory()
in_offs = np.arange(n, dtype=np.uint64) * BLOCK; in_lens = np.full(n, BLOCK, dtype=np.uint32); caps = np.full(n, cap, dtype=np.uint32)
c_lens = np.zeros(n, dtype=np.uint32); d_lens = np.zeros(n, dtype=np.uint32); stt = np.zeros(n * 4, dtype=np.uint64)
offs = np.arange(n, dtype=np.uint64) * np.uint64(45000)
def comp(o):
    assert L.sb_compress_batch_host(h_in.data_ptr(), in_offs.ctypes.data, in_lens.ctypes.data, h_c.data_ptr(), o.ctypes.data, caps.ctypes.data, c_lens.ctypes.data, n, C.byref(err)) == 0
comp(offs)
dense = np.zeros(n, dtype=np.uint64); np.cumsum(c_lens[:-1].astype(np.uint64), out=dense[1:])
def decomp():
    assert L.sb_decompress_batch_host(h_c.data_ptr(), dense.ctypes.data, c_lens.ctypes.data, h_out.data_ptr(), in_offs.ctypes.data, in_lens.ctypes.data, d_lens.ctypes.data, stt.ctypes.data, n, C.byref(err)) == 0
for _ in range(2): comp(dense); decomp()
os.environ["X"]="1"
for name, fn in (("compress_host", lambda: comp(dense)), ("decompress_host", decomp)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(name, "%.1f ms  %.1f GB/s uncompressed" % (dt * 1e3, n * BLOCK / dt / 1e9))
assert torch.equal(h_in, h_out)
d = torch.empty(n * BLOCK, dtype=torch.uint8, device=dev)
for name, a, b in (("H2D", d, h_in), ("D2H", h_out, d)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(name, "%.1f GB/s" % (n * BLOCK / dt / 1e9))
