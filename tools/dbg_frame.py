import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch; torch.cuda.set_device(0)
import gpu_helpers, ctypes as C
from conftest import corpus
s = gpu_helpers.snap(); L = gpu_helpers.lib()
base = corpus("alice29.txt") + corpus("lcet10.txt")
for mult in (1, 8, 64, 300):
    data = (base * mult)[: (len(base) * mult) // 65536 * 65536]
    for ident in (True, False):
        stream, offs, res = gpu_helpers.frame_encode_device_ws(data, ident=ident)
        n = len(stream)
        dev = torch.device("cuda:0")
        t_in = torch.frombuffer(bytearray(stream) + bytearray(16), dtype=torch.uint8).to(dev)
        t_out = torch.zeros(len(data) + 16, dtype=torch.uint8, device=dev)
        t_idx = torch.tensor(offs, dtype=torch.int64, device=dev)
        W = len(offs) - 1
        for maxc in (W + 1, W + 5000):
            sb = L.sb_frame_decode_scratch_bytes(maxc)
            t_scr = torch.zeros(sb + 256, dtype=torch.uint8, device=dev)
            t_res = torch.zeros(8, dtype=torch.int64, device=dev)
            e = s._lib.SbError()
            rc = L.sb_frame_decode_device_ws(t_in.data_ptr(), n, t_out.data_ptr(), len(data), t_idx.data_ptr(), W, 0 if ident else 1,
                                             t_res.data_ptr(), t_scr.data_ptr(), sb + 256, maxc, torch.cuda.current_stream().cuda_stream, C.byref(e))
            torch.cuda.synchronize()
            print("chunks", W, "ident", ident, "maxc", maxc, "rc", rc, "res", t_res[:6].tolist())
