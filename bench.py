#!/usr/bin/env python
"""bench.py -- the headline measurement for the Snappy raw-block hot path.

Workload (BASELINE.json configs[1]): `--blocks` (default 1,048,576) independent
64KB synthetic text blocks per GPU, block i = T[off_i : off_i+65536] with
T = alice29 || asyoulik || lcet10 || plrabn12 and off_i = (i*65521) mod (|T|-65536)
(SURVEY.md 8d). One step = compress every block (K1) then decompress every
compressed block (K2), device resident, in waves that reuse two staging buffers.

metric  : uncompressed GB/s over the compress+decompress round trip
          = 2 * uncompressed_bytes / (t_compress + t_decompress)
value   : device-resident (inputs already in HBM), CUDA events, max over ranks
e2e     : the same round trip through the C ABI with HOST (pinned) buffers,
          H2D/D2H inside the timed region
--impl reference : the reference's CPU implementation of the same path (the
          oracle port -- the Rust crate cannot be built here), all host threads,
          bounded sample per step.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK = 65536
STRIDE = 76544
MUL = 65521
TEXT_FILES = ("alice29.txt", "asyoulik.txt", "lcet10.txt", "plrabn12.txt")
METRIC = "uncompressed GB/s, raw 64KB-block compress+decompress round trip"


def load_text():
    d = os.path.join(ROOT, "tests", "golden", "data")
    return b"".join(open(os.path.join(d, f), "rb").read() for f in TEXT_FILES)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every 200 ms via NVML."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- reference arm
def cpu_roundtrip(orc, text, first, count, threads):
    """Oracle port on host threads: compress `count` blocks, then decompress them."""
    tc, comp_total = orc.bench_compress_mt(text, BLOCK, first, count, MUL, threads)
    # a small fixed set of compressed streams, tiled round-robin like the GPU decode input
    span = len(text) - BLOCK
    streams = [orc.compress(text[((first + i) * MUL) % span:][:BLOCK]) for i in range(min(count, 64))]
    td, dec_total = orc.bench_decompress_mt(streams, count, threads)
    assert dec_total == count * BLOCK
    return tc, td, comp_total


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.lib()
    text = load_text()
    threads = host_threads()
    # probe speed, then size each step to ~6 s of CPU work so W+K steps end within minutes
    tc, td, _ = cpu_roundtrip(orc, text, 0, 64 * threads, threads)
    rate = 64 * threads / (tc + td)
    budget = min(6.0, 150.0 / max(1, args.steps + args.warmup))
    count = max(threads, int(rate * budget))
    for _ in range(args.warmup):
        cpu_roundtrip(orc, text, 0, count, threads)
    ttot_c = ttot_d = 0.0
    for _ in range(args.steps):
        tc, td, _ = cpu_roundtrip(orc, text, 0, count, threads)
        ttot_c += tc
        ttot_d += td
    u = count * BLOCK * args.steps
    val = 2 * u / (ttot_c + ttot_d) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (ttot_c + ttot_d) / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "raw 64KB text blocks (BASELINE configs[1] generator), CPU sample of %d blocks per step" % count,
                   "block_bytes": BLOCK},
        "compress_gbs": u / ttot_c / 1e9, "decompress_gbs": u / ttot_d / 1e9,
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": threads, "kind": "port",
                         "sample": "%d blocks x 64KB per step, oracle C port of rust-snappy (no rustc in image)" % count},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm
def run_ours(args, rank, local_rank, world):
    import torch
    import __graft_entry__ as graft
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        # keep stdout for the one JSON line: NCCL's banner/debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    snap = graft.load_package()
    L = snap._lib.lib()
    err = snap._lib.SbError()
    text = load_text()
    span = len(text) - BLOCK
    blocks = args.blocks
    wave = min(args.wave, blocks)
    nwaves = (blocks + wave - 1) // wave
    first_block = rank * blocks          # weak scaling: every rank owns its own range of blocks

    t_text = torch.frombuffer(bytearray(text), dtype=torch.uint8).to(dev)
    t_in = torch.empty(blocks * BLOCK, dtype=torch.uint8, device=dev)
    t_c = torch.empty(wave * STRIDE, dtype=torch.uint8, device=dev)
    t_out = torch.empty(wave * BLOCK, dtype=torch.uint8, device=dev)
    t_clen = torch.zeros(blocks, dtype=torch.int32, device=dev)
    t_dlen = torch.zeros(wave, dtype=torch.int32, device=dev)
    t_st = torch.zeros(wave * 4, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def ck(rc):
        if rc:
            raise snap.error.from_c(err)

    ck(L.sb_generate_blocks_device(t_text.data_ptr(), len(text), t_in.data_ptr(), BLOCK, BLOCK,
                                   first_block, blocks, MUL, stream, C.byref(err)))
    torch.cuda.synchronize()

    def batch(in_ptr, in_stride, in_len, in_lens_ptr, out_ptr, out_stride, out_cap, lens_ptr, st_ptr, count):
        b = snap._lib.SbBatch()
        b.in_base, b.in_stride, b.in_len_uniform = in_ptr, in_stride, in_len
        if in_lens_ptr:
            b.in_lens = in_lens_ptr
        b.out_base, b.out_stride, b.out_cap_uniform = out_ptr, out_stride, out_cap
        b.out_lens = lens_ptr
        if st_ptr:
            b.statuses = st_ptr
        b.count = count
        return b

    ev = torch.cuda.Event

    def one_step(verify=False):
        """compress + decompress every wave; returns (ms_compress, ms_decompress)."""
        marks = []
        for w in range(nwaves):
            lo = w * wave
            cnt = min(wave, blocks - lo)
            e0, e1, e2 = ev(enable_timing=True), ev(enable_timing=True), ev(enable_timing=True)
            bc = batch(t_in.data_ptr() + lo * BLOCK, BLOCK, BLOCK, 0, t_c.data_ptr(), STRIDE, STRIDE,
                       t_clen.data_ptr() + 4 * lo, 0, cnt)
            bd = batch(t_c.data_ptr(), STRIDE, 0, t_clen.data_ptr() + 4 * lo, t_out.data_ptr(), BLOCK, BLOCK,
                       t_dlen.data_ptr(), t_st.data_ptr(), cnt)
            e0.record()
            ck(L.sb_compress_batch_device(C.byref(bc), stream, C.byref(err)))
            e1.record()
            ck(L.sb_decompress_batch_device(C.byref(bd), stream, C.byref(err)))
            e2.record()
            marks.append((e0, e1, e2))
            if verify:
                torch.cuda.synchronize()
                assert torch.equal(t_in[lo * BLOCK:(lo + cnt) * BLOCK], t_out[:cnt * BLOCK]), "round trip mismatch"
                assert int(t_st.view(wave, 4)[:cnt, 0].abs().sum()) == 0, "decode status != Ok"
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b, _ in marks), sum(b.elapsed_time(c) for _, b, c in marks)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (first warm-up step also verifies the round trip bit-exactly on device)
    for i in range(args.warmup):
        one_step(verify=(i == 0))
    comp_bytes = int(t_clen.to(torch.int64).sum().item())
    # sampled bit-exact check of compressed bytes against the oracle (outside the timed region)
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import oracle as orc
        lo = (nwaves - 1) * wave
        cnt = blocks - lo
        idx = sorted(set([0, cnt - 1] + [(k * 7919) % cnt for k in range(args.parity_samples)]))
        clen = t_clen[lo:lo + cnt].cpu().numpy()
        ok_n = 0
        for i in idx:
            got = bytes(t_c[i * STRIDE:i * STRIDE + int(clen[i])].cpu().numpy())
            off = ((first_block + lo + i) * MUL) % span
            ok_n += int(got == orc.compress(text[off:off + BLOCK]))
        parity = {"blocks_compared": len(idx), "blocks_equal": ok_n}
        assert ok_n == len(idx), "compressed bytes differ from the oracle"

    sampler = ClockSampler(local_rank)
    barrier()
    launches0 = L.sb_launch_count()
    sampler.start()
    t0 = time.perf_counter()
    ms_c = ms_d = 0.0
    for _ in range(args.steps):
        c_ms, d_ms = one_step()
        ms_c += c_ms
        ms_d += d_ms
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = L.sb_launch_count() - launches0

    # max over ranks of the device-timed step
    tot = torch.tensor([ms_c + ms_d, ms_c, ms_d], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        sizes = torch.tensor([comp_bytes], dtype=torch.int64, device=dev)
        gathered = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(gathered, sizes)          # per-rank compressed totals -> global stream offsets
        comp_all = int(sum(int(g.item()) for g in gathered))
    else:
        comp_all = comp_bytes
    ms_tot, ms_cmax, ms_dmax = [float(x) for x in tot.tolist()]
    u_rank = blocks * BLOCK
    u_all = u_rank * world

    # ---------------- e2e: host buffers through the C ABI (H2D/D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, snap, L, torch, dev, t_in, t_clen, rank, world)

    if rank != 0:
        return
    peak, peak_src = measured_peak()
    value = 2 * u_all * args.steps / (ms_tot / 1e3) / 1e9
    k1_bytes = (u_rank + comp_bytes) * args.steps          # algorithmic bytes moved by K1 launches
    k1_achieved = k1_bytes / (ms_c / 1e3) / 1e9
    k2_achieved = k1_bytes / (ms_d / 1e3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_block") * min(wave, blocks)   # per launch
        except Exception:  # noqa: BLE001
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_tot / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "batched raw block compress+decompress: %d x 64KB synthetic text blocks per GPU (BASELINE configs[1])" % blocks,
                   "blocks_per_gpu": blocks, "block_bytes": BLOCK, "wave_blocks": wave, "ratio": comp_bytes / u_rank,
                   "l2": "inputs larger than L2 (%.1f GiB per GPU per pass)" % (u_rank / 2**30), "parity": parity,
                   "wall_s_timed_region": wall,
                   # which build/knobs produced the line (A/B runs of experimental libraries set these)
                   "lib": os.path.basename(os.environ.get("SNAPB200_LIB", "libsnapb200.so")),
                   "k1_ng_env": os.environ.get("SNAPB200_K1_NG")},
        "compress_gbs": u_all * args.steps / (ms_cmax / 1e3) / 1e9,
        "decompress_gbs": u_all * args.steps / (ms_dmax / 1e3) / 1e9,
        "roofline": {"bound": "hbm", "kernel": "k1_m7_kernel (K1 compress)", "achieved": k1_achieved, "peak": peak, "unit": "GB/s",
                     "frac": k1_achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": (u_rank + comp_bytes) / nwaves,
                     "k2_decompress_kernel": {"achieved": k2_achieved, "frac": k2_achieved / peak}},
        "clocks": clocks, "gpu_launches": int(launches), "compressed_bytes_all_ranks": comp_all,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        threads = host_threads()
        tc, td, _ = cpu_roundtrip(orc, text, 0, 32 * threads, threads)
        count = max(threads, int(32 * threads / (tc + td) * 12.0))
        tc, td, _ = cpu_roundtrip(orc, text, 0, count, threads)
        line["cpu_baseline"] = {"value": 2 * count * BLOCK / (tc + td) / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
                                "sample": "%d of the same 64KB text blocks, compress+decompress, oracle C port on all host threads" % count,
                                "compress_gbs": count * BLOCK / tc / 1e9, "decompress_gbs": count * BLOCK / td / 1e9}
    print(json.dumps(line), flush=True)


def run_e2e(args, snap, L, torch, dev, t_in, t_clen, rank, world):
    """Round trip through sb_compress_batch_host / sb_decompress_batch_host with pinned host buffers."""
    import numpy as np
    avail = 0
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                avail = int(ln.split()[1]) * 1024
    except OSError:
        pass
    n = min(args.e2e_blocks, args.blocks)
    while n > 1024 and avail and n * BLOCK * 3.2 * max(1, world) > 0.5 * avail:
        n //= 2
    err = snap._lib.SbError()
    h_in = torch.empty(n * BLOCK, dtype=torch.uint8).pin_memory()
    h_in.copy_(t_in[:n * BLOCK])
    cap = int(L.sb_max_compress_len(BLOCK))
    known = t_clen[:n].cpu().numpy().astype(np.uint64)      # sizes from the device-resident pass
    h_c = torch.empty(int(known.sum()) + cap, dtype=torch.uint8).pin_memory()   # dense compressed stream
    h_out = torch.empty(n * BLOCK, dtype=torch.uint8).pin_memory()
    in_offs = np.arange(n, dtype=np.uint64) * BLOCK
    in_lens = np.full(n, BLOCK, dtype=np.uint32)
    caps = np.full(n, cap, dtype=np.uint32)
    c_lens = np.zeros(n, dtype=np.uint32)
    d_lens = np.zeros(n, dtype=np.uint32)
    st = np.zeros(n * 4, dtype=np.uint64)
    c_offs = np.zeros(n, dtype=np.uint64)

    def step():
        # back-to-back destination offsets -> the library gathers each wave on the device and
        # drains it with one D2H copy
        rc = L.sb_compress_batch_host(h_in.data_ptr(), in_offs.ctypes.data, in_lens.ctypes.data, h_c.data_ptr(),
                                      dense_offs.ctypes.data, caps.ctypes.data, c_lens.ctypes.data, n, C.byref(err))
        if rc:
            raise snap.error.from_c(err)
        np.cumsum(c_lens[:-1], dtype=np.uint64, out=c_offs[1:])
        rc = L.sb_decompress_batch_host(h_c.data_ptr(), c_offs.ctypes.data, c_lens.ctypes.data, h_out.data_ptr(),
                                        in_offs.ctypes.data, in_lens.ctypes.data, d_lens.ctypes.data, st.ctypes.data, n,
                                        C.byref(err))
        if rc:
            raise snap.error.from_c(err)

    # dense destinations: unit k lands right after unit k-1 (sizes known from the device-resident pass)
    dense_offs = np.zeros(n, dtype=np.uint64)
    np.cumsum(known[:-1], dtype=np.uint64, out=dense_offs[1:])
    for _ in range(max(1, args.warmup - 1)):
        step()
    assert bool((d_lens == BLOCK).all()) and torch.equal(h_in, h_out), "e2e round trip mismatch"
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    cbytes = int(c_lens.astype(np.uint64).sum())
    return {"value": 2 * n * BLOCK * world * args.steps / dt / 1e9, "unit": "GB/s",
            "h2d_bytes_per_step": n * BLOCK + cbytes, "d2h_bytes_per_step": cbytes + n * BLOCK,
            "blocks_per_gpu": n, "api": "sb_compress_batch_host + sb_decompress_batch_host (pinned host buffers)",
            "ms_per_step": 1e3 * dt / args.steps}


def run_urls(args, local_rank):
    """BASELINE configs[2]: urls.10K cut into 11 blocks, each compressed independently, tiled
    round-robin (compressed bytes physically replicated in HBM) and decoded by K2."""
    import numpy as np
    import torch
    import __graft_entry__ as graft
    from oracle import oracle as orc
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    snap = graft.load_package()
    L = snap._lib.lib()
    err = snap._lib.SbError()
    data = open(os.path.join(ROOT, "tests", "golden", "data", "urls.10K"), "rb").read()
    blocks = [data[i:i + BLOCK] for i in range(0, len(data), BLOCK)]
    comp = [orc.compress(b) for b in blocks]                  # input preparation, outside the timed region
    reps = int(args.urls_gib * (1 << 30)) // len(data)
    n = reps * len(blocks)
    clen = np.array([len(c) for c in comp], dtype=np.int64)
    tile = int(clen.sum())
    src = torch.frombuffer(bytearray(b"".join(comp)), dtype=torch.uint8).to(dev)
    t_c = src.repeat(reps)                                    # physical tiling of the compressed streams
    starts = np.concatenate([[0], np.cumsum(clen)[:-1]])
    base = (np.arange(reps, dtype=np.int64) * tile)[:, None] + starts[None, :]
    in_ptrs = torch.from_numpy(base.reshape(-1) + t_c.data_ptr()).to(dev)
    in_lens = torch.from_numpy(np.tile(clen, reps).astype(np.int32)).to(dev)
    t_out = torch.empty(n * BLOCK, dtype=torch.uint8, device=dev)
    t_dlen = torch.zeros(n, dtype=torch.int32, device=dev)
    t_st = torch.zeros(n * 4, dtype=torch.int64, device=dev)
    b = snap._lib.SbBatch()
    b.in_ptrs, b.in_lens = in_ptrs.data_ptr(), in_lens.data_ptr()
    b.out_base, b.out_stride, b.out_cap_uniform = t_out.data_ptr(), BLOCK, BLOCK
    b.out_lens, b.statuses, b.count = t_dlen.data_ptr(), t_st.data_ptr(), n
    stream = torch.cuda.current_stream().cuda_stream
    ev = torch.cuda.Event

    def step():
        e0, e1 = ev(enable_timing=True), ev(enable_timing=True)
        e0.record()
        if L.sb_decompress_batch_device(C.byref(b), stream, C.byref(err)):
            raise snap.error.from_c(err)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    for _ in range(max(3, args.warmup)):
        step()
    assert int(t_st.view(n, 4)[:, 0].abs().sum()) == 0
    view = t_out.view(reps, len(blocks), BLOCK)
    for k, blk in enumerate(blocks):                          # every tile decodes to the original bytes
        want = torch.frombuffer(bytearray(blk), dtype=torch.uint8).to(dev)
        assert bool((view[:, k, :len(blk)] == want).all())
    ms = sum(step() for _ in range(args.steps)) / args.steps
    u, c = reps * len(data), reps * tile
    peak, peak_src = measured_peak()
    print(json.dumps({
        "metric": "uncompressed GB/s, batched raw block decompress", "value": u / (ms / 1e3) / 1e9, "unit": "GB/s",
        "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True,
        "dtype": "u8", "data": "data/urls.10K tiled", "side_measurement": True,
        "config": {"workload": "batched raw block decompress: data/urls.10K tiled to %.1f GiB (BASELINE configs[2])" % (u / 2**30),
                   "streams": n, "compressed_bytes": c, "ratio": c / u, "parity": "every stream equals its source block"},
        "roofline": {"bound": "hbm", "kernel": "k2_decompress_kernel", "achieved": (u + c) / (ms / 1e3) / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": (u + c) / (ms / 1e3) / 1e9 / peak, "peak_source": peak_src, "traffic": None},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--blocks", type=int, default=1 << 20, help="64KB blocks per GPU (BASELINE configs[1]: 1M)")
    ap.add_argument("--wave", type=int, default=1 << 17, help="blocks per kernel launch")
    ap.add_argument("--e2e-blocks", type=int, default=1 << 18)
    ap.add_argument("--parity-samples", type=int, default=48)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="text-roundtrip", choices=["text-roundtrip", "urls-decompress"],
                    help="urls-decompress = BASELINE configs[2] (data/urls.10K tiled), decompress only; a side measurement")
    ap.add_argument("--urls-gib", type=float, default=64.0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "urls-decompress":
        run_urls(args, local_rank)
        return
    run_ours(args, rank, local_rank, world)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
