"""Per-corpus-file device-resident throughput (the analogue of the reference README's table, README.md:126-163):
blocks of 64KB cut from each data/ file at offsets (i*65521) mod (len-65536), K1 then K2, round trip verified."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as graft
BLOCK, STRIDE, MUL = 65536, 76544, 65521
snap = graft.load_package(); L = snap._lib.lib(); err = snap._lib.SbError()
torch.cuda.set_device(0); dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
out = {}
for name in ["html", "urls.10K", "fireworks.jpeg", "paper-100k.pdf", "html_x_4", "alice29.txt", "asyoulik.txt", "lcet10.txt", "plrabn12.txt", "geo.protodata", "kppkn.gtb"]:
    data = open(os.path.join(ROOT, "tests", "golden", "data", name), "rb").read()
    t_text = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
    t_in = torch.empty(n * BLOCK, dtype=torch.uint8, device=dev)
    t_c = torch.empty(n * STRIDE, dtype=torch.uint8, device=dev)
    t_out = torch.zeros(n * BLOCK, dtype=torch.uint8, device=dev)
    cl = torch.zeros(n, dtype=torch.int32, device=dev); dl = torch.zeros(n, dtype=torch.int32, device=dev)
    stt = torch.zeros(n * 4, dtype=torch.int64, device=dev)
    assert L.sb_generate_blocks_device(t_text.data_ptr(), len(data), t_in.data_ptr(), BLOCK, BLOCK, 0, n, MUL, st, C.byref(err)) == 0
    bc = snap._lib.SbBatch(); bc.in_base, bc.in_stride, bc.in_len_uniform = t_in.data_ptr(), BLOCK, BLOCK
    bc.out_base, bc.out_stride, bc.out_cap_uniform, bc.out_lens, bc.count = t_c.data_ptr(), STRIDE, STRIDE, cl.data_ptr(), n
    bd = snap._lib.SbBatch(); bd.in_base, bd.in_stride, bd.in_lens = t_c.data_ptr(), STRIDE, cl.data_ptr()
    bd.out_base, bd.out_stride, bd.out_cap_uniform, bd.out_lens, bd.statuses, bd.count = t_out.data_ptr(), BLOCK, BLOCK, dl.data_ptr(), stt.data_ptr(), n
    def go():
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); assert L.sb_compress_batch_device(C.byref(bc), st, C.byref(err)) == 0
        e[1].record(); assert L.sb_decompress_batch_device(C.byref(bd), st, C.byref(err)) == 0
        e[2].record(); torch.cuda.synchronize()
        return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    go()
    assert torch.equal(t_in, t_out) and int(stt.view(n, 4)[:, 0].abs().sum()) == 0
    c_ms, d_ms = go()
    u = n * BLOCK
    out[name] = {"ratio": round(float(cl.sum().item()) / u, 4), "compress_gbs": round(u / c_ms / 1e6, 2), "decompress_gbs": round(u / d_ms / 1e6, 2)}
    print(name, out[name], flush=True)
print(json.dumps(out))
