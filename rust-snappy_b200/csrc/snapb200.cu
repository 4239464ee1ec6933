// snapb200.cu -- libsnapb200.so: sm_100a kernels + the C ABI of include/snapb200.h.
// Built by __graft_entry__.build() with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
// There is no CPU execution path in this library: every compute entry point
// launches the kernels below and fails with SB_E_NO_DEVICE when it cannot.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>
#include <time.h>

#include "k1_compress.cuh"
#include "k2_decompress.cuh"
#include "k3_crc32c.cuh"
#include "k4_frame.cuh"

namespace {

// ------------------------------------------------------------------ kernels
// K1 variants: window in shared memory (S) or read in place from global/L2 (G) x parser warps per block
__global__ void __launch_bounds__(64) k1_s1_kernel(sb_batch b, uint32_t flags) { sbk::k1_compress_body<false, 1>(b, flags); }
__global__ void __launch_bounds__(96) k1_s2_kernel(sb_batch b, uint32_t flags) { sbk::k1_compress_body<false, 2>(b, flags); }
__global__ void __launch_bounds__(128) k1_s3_kernel(sb_batch b, uint32_t flags) { sbk::k1_compress_body<false, 3>(b, flags); }
__global__ void __launch_bounds__(160) k1_s4_kernel(sb_batch b, uint32_t flags) { sbk::k1_compress_body<false, 4>(b, flags); }
__global__ void __launch_bounds__(64) k1_g1_kernel(sb_batch b, uint32_t flags) { sbk::k1_compress_body<true, 1>(b, flags); }
__global__ void __launch_bounds__(96) k1_g2_kernel(sb_batch b, uint32_t flags) { sbk::k1_compress_body<true, 2>(b, flags); }
__global__ void __launch_bounds__(128) k1_g3_kernel(sb_batch b, uint32_t flags) { sbk::k1_compress_body<true, 3>(b, flags); }
// one CTA per SM: 7 parser/emitter pairs with their tables in shared memory + NG pairs with
// their tables in an L2-resident scratch; rings in global scratch; units taken from `work`
template <int NG>
__global__ void __launch_bounds__((7 + NG) * 64, 1)
k1_m7_kernel(sb_batch b, uint32_t flags, uint64_t* rings, uint16_t* gtables, uint32_t* work) {
    sbk::k1_compress_body_multi<7, NG>(b, flags, rings, gtables, work);
}
// second-generation parser (k1_exact.cuh): 6 shared-memory tables + per-chain byte/info rings, NG chains with L2 tables
template <int NG>
__global__ void __launch_bounds__((6 + NG) * 64, 1)
k1_x_kernel(sb_batch b, uint32_t flags, uint64_t* rings, uint16_t* gtables, uint32_t* work) {
    sbk::k1_compress_body_multi<6, NG, true>(b, flags, rings, gtables, work);
}
const int K1X_MAX_NG = 10;
const int K1_MAX_NG = 7;
const size_t K1_M7_SMEM = 7 * sbk::K1_TABLE_BYTES + (7 + K1_MAX_NG) * 64;
__global__ void __launch_bounds__(128) k2_decompress_kernel(sb_batch b) { sbk::k2_decompress_body(b); }
__global__ void __launch_bounds__(256) k3_crc_kernel(sb_batch b) { sbk::k3_crc_body(b); }
__global__ void __launch_bounds__(256) k4_sizes_kernel(sbk::FramePlan p) { sbk::k4_sizes_body(p); }
__global__ void __launch_bounds__(1024) k4_scan_kernel(sbk::FramePlan p) { sbk::k4_scan_body(p); }
__global__ void __launch_bounds__(256) k4_gather_kernel(sbk::FramePlan p) { sbk::k4_gather_body(p); }
__global__ void __launch_bounds__(256) k5_copy_units_kernel(sb_batch b) { sbk::k5_copy_units_body(b); }
__global__ void __launch_bounds__(256) k6_generate_kernel(sbk::GenPlan g) { sbk::k6_generate_body(g); }

std::atomic<uint64_t> g_launches{0};
const int K1_DEFAULT_NP = 1;
const int K2_DEFAULT_CTAS_PER_SM = 16;
const int K1_DEFAULT_GW = 1;
const int K1_DEFAULT_MULTI = 1;
const int K1_DEFAULT_NG = 5;

int fail(sb_error* e, uint32_t code, uint64_t a = 0, uint64_t b = 0, uint64_t c = 0) {
    if (e) { e->code = code; e->_pad = 0; e->a = a; e->b = b; e->c = c; }
    return (int)code;
}
void ok(sb_error* e) { if (e) { e->code = 0; e->_pad = 0; e->a = e->b = e->c = 0; } }

#define CK(call)                                                                   \
    do {                                                                           \
        cudaError_t _e = (call);                                                   \
        if (_e != cudaSuccess) {                                                   \
            if (getenv("SNAPB200_DEBUG"))                                          \
                fprintf(stderr, "snapb200: %s -> %s (%s:%d)\n", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return fail(err, (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver || \
                              _e == cudaErrorNoKernelImageForDevice) ? SB_E_NO_DEVICE : SB_E_CUDA, (uint64_t)_e); \
        }                                                                          \
    } while (0)

// ------------------------------------------------------------- device state
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t need(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return e; }
        size_t want = n + n / 8 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, n); want = n; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <class T> T* as() { return (T*)p; }
};

struct Ctx {
    int dev = -1, sms = 0;
    bool ready = false;
    cudaStream_t s_compute = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    DevBuf rings, gtables, work;   // K1 scratch: event rings, L2-resident tables, unit counter
    cudaEvent_t k1_done = nullptr; // K1 launches share that scratch: each waits for the previous one, whatever its stream
    std::mutex k1_mu;
    DevBuf in[2], slots[2], compact[2], lens[2], csize[2], offs[2], crcs[2], status[2], ptrs_in[2], ptrs_out[2], caps[2];
    void* pinned[4] = {nullptr, nullptr, nullptr, nullptr}; size_t pinned_cap[4] = {0, 0, 0, 0};   // pinned staging: [0,1] descriptors in, [2,3] results out
    std::mutex mu;
};
Ctx g_ctx[16];

int get_ctx(Ctx** out, sb_error* err) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) return fail(err, SB_E_NO_DEVICE);
    Ctx& c = g_ctx[dev];
    if (!c.ready) {
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, dev));
        c.dev = dev; c.sms = prop.multiProcessorCount;
        CK(cudaFuncSetAttribute(k1_s1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbk::K1_SMEM_BYTES));
        CK(cudaFuncSetAttribute(k1_s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbk::K1_SMEM_BYTES));
        CK(cudaFuncSetAttribute(k1_s3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbk::K1_SMEM_BYTES));
        CK(cudaFuncSetAttribute(k1_s4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbk::K1_SMEM_BYTES));
        CK(cudaFuncSetAttribute(k1_g1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbk::K1_SMEM_BYTES_GW));
        CK(cudaFuncSetAttribute(k1_g2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbk::K1_SMEM_BYTES_GW));
        CK(cudaFuncSetAttribute(k1_g3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbk::K1_SMEM_BYTES_GW));
        CK(cudaFuncSetAttribute(k1_m7_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1_M7_SMEM));
        CK(cudaFuncSetAttribute(k1_m7_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1_M7_SMEM));
        CK(cudaFuncSetAttribute(k1_m7_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1_M7_SMEM));
        CK(cudaFuncSetAttribute(k1_x_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbk::k1_multi_smem(6, 0, true)));
        CK(cudaFuncSetAttribute(k1_x_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbk::k1_multi_smem(6, 4, true)));
        CK(cudaFuncSetAttribute(k1_x_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbk::k1_multi_smem(6, 6, true)));
        CK(cudaFuncSetAttribute(k1_x_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbk::k1_multi_smem(6, 8, true)));
        CK(cudaFuncSetAttribute(k1_x_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbk::k1_multi_smem(6, 10, true)));
        CK(c.rings.need((size_t)c.sms * 16 * sbk::K1_RING_GW * 8));
        CK(c.gtables.need((size_t)c.sms * K1X_MAX_NG * sbk::K1_TABLE_BYTES));
        CK(c.work.need(256));
        CK(cudaEventCreateWithFlags(&c.k1_done, cudaEventDisableTiming));
        CK(cudaStreamCreateWithFlags(&c.s_compute, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c.s_h2d, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c.s_d2h, cudaStreamNonBlocking));
        c.ready = true;
    }
    *out = &c;
    return 0;
}

// Descriptor arrays (pointers, lengths) are staged through pinned memory: an async copy from
// pageable memory would not overlap with the running kernel.
int need_pinned(Ctx& c, int slot, size_t n, sb_error* err) {
    if (n <= c.pinned_cap[slot]) return 0;
    if (c.pinned[slot]) { CK(cudaFreeHost(c.pinned[slot])); c.pinned[slot] = nullptr; c.pinned_cap[slot] = 0; }
    n += n / 4 + 4096;
    CK(cudaHostAlloc(&c.pinned[slot], n, cudaHostAllocDefault));
    c.pinned_cap[slot] = n;
    return 0;
}

// ------------------------------------------------------------ launch helpers
int launch_k1(Ctx& c, const sb_batch& b, uint32_t flags, cudaStream_t st, sb_error* err) {
    if (b.count == 0) return 0;
    // K1 variant: SNAPB200_K1_GW=1 (default) reads the window in place from global/L2 (6 CTAs/SM instead of 2),
    // SNAPB200_K1_NP = parser warps per block (pipelined over windows).
    static const int gw = getenv("SNAPB200_K1_GW") ? atoi(getenv("SNAPB200_K1_GW")) : K1_DEFAULT_GW;
    static const int np = getenv("SNAPB200_K1_NP") ? atoi(getenv("SNAPB200_K1_NP")) : K1_DEFAULT_NP;
    unsigned grid = (unsigned)((gw ? 6 : 2) * c.sms);
    if (grid > b.count) grid = b.count;
    const size_t sm = gw ? sbk::K1_SMEM_BYTES_GW : sbk::K1_SMEM_BYTES;
    static const int multi = getenv("SNAPB200_K1_MULTI") ? atoi(getenv("SNAPB200_K1_MULTI")) : K1_DEFAULT_MULTI;
    static const int xmode = getenv("SNAPB200_K1_X") ? atoi(getenv("SNAPB200_K1_X")) : 1;
    if (xmode) {
        // SNAPB200_K1_NG = chains per SM with L2-resident tables next to the 6 shared-memory ones (0..10)
        static const int ng_env = getenv("SNAPB200_K1_NG") ? atoi(getenv("SNAPB200_K1_NG")) : 6;
        const unsigned ng = ng_env < 0 ? 0 : ng_env > K1X_MAX_NG ? K1X_MAX_NG : (unsigned)ng_env;
        unsigned chains = (unsigned)(((uint64_t)b.count + c.sms - 1) / c.sms);
        if (chains > 6 + ng) chains = 6 + ng;
        unsigned mg = (unsigned)c.sms;
        if (mg > b.count) mg = b.count;
        std::lock_guard<std::mutex> k1lk(c.k1_mu);
        CK(cudaStreamWaitEvent(st, c.k1_done, 0));
        CK(cudaMemsetAsync(c.work.p, 0, 4, st));
        uint64_t* rg = c.rings.as<uint64_t>(); uint16_t* gt = c.gtables.as<uint16_t>(); uint32_t* wk = c.work.as<uint32_t>();
        // the template argument bounds the chain count (launch bounds / register cap, scratch strides)
        if (ng > 8) k1_x_kernel<10><<<mg, chains * 64, sbk::k1_multi_smem(6, 10, true), st>>>(b, flags, rg, gt, wk);
        else if (ng > 6) k1_x_kernel<8><<<mg, chains * 64, sbk::k1_multi_smem(6, 8, true), st>>>(b, flags, rg, gt, wk);
        else if (ng > 4) k1_x_kernel<6><<<mg, chains * 64, sbk::k1_multi_smem(6, 6, true), st>>>(b, flags, rg, gt, wk);
        else if (ng > 0) k1_x_kernel<4><<<mg, chains * 64, sbk::k1_multi_smem(6, 4, true), st>>>(b, flags, rg, gt, wk);
        else k1_x_kernel<0><<<mg, chains * 64, sbk::k1_multi_smem(6, 0, true), st>>>(b, flags, rg, gt, wk);
        CK(cudaGetLastError());
        CK(cudaEventRecord(c.k1_done, st));
    } else if (multi) {
        // SNAPB200_K1_NG = extra chains per SM with L2-resident tables (0..7)
        static const int ng_env = getenv("SNAPB200_K1_NG") ? atoi(getenv("SNAPB200_K1_NG")) : K1_DEFAULT_NG;
        const unsigned ng = ng_env < 0 ? 0 : ng_env > K1_MAX_NG ? K1_MAX_NG : (unsigned)ng_env;
        // small batches spread over the SMs first (one shared-memory-table chain per SM is the fastest a block can
        // run); only batches with more units than that stack chains on an SM, L2-table chains last
        unsigned chains = (unsigned)(((uint64_t)b.count + c.sms - 1) / c.sms);
        if (chains > 7 + ng) chains = 7 + ng;
        unsigned mg = (unsigned)c.sms;
        if (mg > b.count) mg = b.count;
        std::lock_guard<std::mutex> k1lk(c.k1_mu);
        CK(cudaStreamWaitEvent(st, c.k1_done, 0));
        CK(cudaMemsetAsync(c.work.p, 0, 4, st));
        // the template argument only bounds the chain count (launch bounds / register cap, scratch strides)
        if (ng > 5) k1_m7_kernel<7><<<mg, chains * 64, K1_M7_SMEM, st>>>(b, flags, c.rings.as<uint64_t>(), c.gtables.as<uint16_t>(), c.work.as<uint32_t>());
        else if (ng > 0) k1_m7_kernel<5><<<mg, chains * 64, K1_M7_SMEM, st>>>(b, flags, c.rings.as<uint64_t>(), c.gtables.as<uint16_t>(), c.work.as<uint32_t>());
        else k1_m7_kernel<0><<<mg, chains * 64, K1_M7_SMEM, st>>>(b, flags, c.rings.as<uint64_t>(), c.gtables.as<uint16_t>(), c.work.as<uint32_t>());
        CK(cudaGetLastError());
        CK(cudaEventRecord(c.k1_done, st));
    } else if (gw) {
        if (np <= 1) k1_g1_kernel<<<grid, 64, sm, st>>>(b, flags);
        else if (np == 2) k1_g2_kernel<<<grid, 96, sm, st>>>(b, flags);
        else k1_g3_kernel<<<grid, 128, sm, st>>>(b, flags);
    } else {
        if (np <= 1) k1_s1_kernel<<<grid, 64, sm, st>>>(b, flags);
        else if (np == 2) k1_s2_kernel<<<grid, 96, sm, st>>>(b, flags);
        else if (np == 3) k1_s3_kernel<<<grid, 128, sm, st>>>(b, flags);
        else k1_s4_kernel<<<grid, 160, sm, st>>>(b, flags);
    }
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}
int launch_k2(Ctx& c, const sb_batch& b, cudaStream_t st, sb_error* err) {
    if (b.count == 0) return 0;
    const unsigned wpb = 4;
    uint64_t blocks = ((uint64_t)b.count + wpb - 1) / wpb;
    // resident CTAs per SM: each warp keeps a 64KB output history alive, and copy sources are
    // re-read from it -- too many streams in flight and the history falls out of the 126MB L2
    static const int per_sm = getenv("SNAPB200_K2_CTAS") ? atoi(getenv("SNAPB200_K2_CTAS")) : K2_DEFAULT_CTAS_PER_SM;
    unsigned grid = (unsigned)(per_sm * c.sms);
    if (grid > blocks) grid = (unsigned)blocks;
    k2_decompress_kernel<<<grid, 32 * wpb, wpb * sbk::K2_SMEM_PER_WARP, st>>>(b);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}
int launch_k3(Ctx& c, const sb_batch& b, cudaStream_t st, sb_error* err) {
    if (b.count == 0) return 0;
    const unsigned wpb = 8;
    uint64_t blocks = ((uint64_t)b.count + wpb - 1) / wpb;
    unsigned grid = (unsigned)(8 * c.sms);
    if (grid > blocks) grid = (unsigned)blocks;
    k3_crc_kernel<<<grid, 32 * wpb, sbk::K3_TABLE_BYTES, st>>>(b);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}
int launch_copy_units(Ctx& c, const sb_batch& b, cudaStream_t st, sb_error* err) {
    if (b.count == 0) return 0;
    uint64_t blocks = ((uint64_t)b.count + 7) / 8;
    unsigned grid = (unsigned)(8 * c.sms);
    if (grid > blocks) grid = (unsigned)blocks;
    k5_copy_units_kernel<<<grid, 256, 0, st>>>(b);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}
// sizes -> scan -> gather over a FramePlan whose slots/clens(/crcs) are filled
int launch_assemble(Ctx& c, const sbk::FramePlan& p, cudaStream_t st, sb_error* err) {
    if (p.nchunks == 0) return 0;
    k4_sizes_kernel<<<(p.nchunks + 255) / 256, 256, 0, st>>>(p);
    k4_scan_kernel<<<1, 1024, 1024 * sizeof(uint64_t), st>>>(p);
    uint64_t blocks = ((uint64_t)p.nchunks + 7) / 8;
    unsigned grid = (unsigned)(8 * c.sms);
    if (grid > blocks) grid = (unsigned)blocks;
    k4_gather_kernel<<<grid, 256, 0, st>>>(p);
    g_launches += 3;
    CK(cudaGetLastError());
    return 0;
}

size_t put_varint(uint8_t* dst, uint64_t v) {   // reference src/bytes.rs:61-70
    size_t i = 0;
    while (v >= 0x80) { dst[i++] = (uint8_t)v | 0x80; v >>= 7; }
    dst[i++] = (uint8_t)v;
    return i;
}
// reference src/bytes.rs:73-90 (checked_shl fails only when shift >= 64)
size_t get_varint(const uint8_t* p, size_t n, uint64_t* out) {
    uint64_t v = 0;
    unsigned shift = 0;
    for (size_t i = 0; i < n; i++) {
        if (shift >= 64) return 0;
        uint8_t b = p[i];
        if (b < 0x80) { *out = v | ((uint64_t)b << shift); return i + 1; }
        v |= (uint64_t)(b & 0x7F) << shift;
        shift += 7;
    }
    return 0;
}

const uint64_t SB_MAX_INPUT = 0xFFFFFFFFull;
const uint32_t SB_MAX_BLOCK = 65536;
const uint32_t SB_MAX_CBLOCK = 76490;   // reference src/frame.rs:12

// Device-resident compress of one logical stream of n bytes at d_in into the
// final layout at d_out: frame=0 -> raw stream (varint + blocks), frame=1 ->
// frame chunks (optionally preceded by the stream identifier). wave buffers b=0.
int compress_stream_device(Ctx& c, const uint8_t* d_in, uint64_t n, uint8_t* d_out, int frame, int ident,
                           uint64_t* total_out, cudaStream_t st, sb_error* err) {
    const uint64_t nchunks64 = (n + SB_MAX_BLOCK - 1) / SB_MAX_BLOCK;
    if (nchunks64 > 0xFFFFFFFFull) return fail(err, SB_TOO_BIG, n, SB_MAX_INPUT);
    const uint32_t nchunks = (uint32_t)nchunks64;
    uint8_t head[16];
    size_t head_len = 0;
    if (frame) { if (ident && n) { memcpy(head, "\xff\x06\x00\x00sNaPpY", 10); head_len = 10; } }
    else head_len = put_varint(head, n);
    if (n == 0) {
        if (head_len) CK(cudaMemcpyAsync(d_out, head, head_len, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        *total_out = head_len;
        return 0;
    }
    CK(c.slots[0].need((size_t)nchunks * sbk::kSlotStride));
    CK(c.lens[0].need((size_t)nchunks * 4 + 4));
    CK(c.csize[0].need((size_t)nchunks * 4 + 4));
    CK(c.offs[0].need(((size_t)nchunks + 1) * 8));
    CK(c.crcs[0].need((size_t)nchunks * 4 + 4));
    CK(c.caps[0].need((size_t)nchunks * 4 + 4));
    // per-chunk input lengths: all 65536 except the last
    {
        std::vector<uint32_t> lens(nchunks, SB_MAX_BLOCK);
        lens[nchunks - 1] = (uint32_t)(n - (uint64_t)(nchunks - 1) * SB_MAX_BLOCK);
        CK(cudaMemcpyAsync(c.caps[0].p, lens.data(), (size_t)nchunks * 4, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));   // `lens` is a temporary
    }
    sb_batch b;
    memset(&b, 0, sizeof b);
    b.in_base = d_in; b.in_stride = SB_MAX_BLOCK; b.in_lens = c.caps[0].as<uint32_t>();
    b.out_base = c.slots[0].as<uint8_t>(); b.out_stride = sbk::kSlotStride; b.out_cap_uniform = sbk::kSlotStride;
    b.out_lens = c.lens[0].as<uint32_t>(); b.count = nchunks;
    int rc = launch_k1(c, b, frame ? 1u : 0u, st, err);   // frame chunks carry their own varint
    if (rc) return rc;
    if (frame) {
        sb_batch cb = b;
        cb.out_lens = c.crcs[0].as<uint32_t>();
        rc = launch_k3(c, cb, st, err);
        if (rc) return rc;
    }
    sbk::FramePlan p;
    p.in = d_in; p.n = n; p.slots = c.slots[0].as<uint8_t>(); p.clens = c.lens[0].as<uint32_t>();
    p.crcs = c.crcs[0].as<uint32_t>(); p.nchunks = nchunks; p.frame = frame ? 1u : 0u; p.base = head_len;
    p.csize = c.csize[0].as<uint32_t>(); p.offs = c.offs[0].as<uint64_t>(); p.out = d_out;
    rc = launch_assemble(c, p, st, err);
    if (rc) return rc;
    if (head_len) CK(cudaMemcpyAsync(d_out, head, head_len, cudaMemcpyHostToDevice, st));
    uint64_t total = 0;
    CK(cudaMemcpyAsync(&total, c.offs[0].as<uint64_t>() + nchunks, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *total_out = total;
    return 0;
}

}  // namespace

// =========================================================================
extern "C" {

const char* sb_version(void) { return "snapb200 0.1 (sm_100a)"; }
uint64_t sb_launch_count(void) { return g_launches.load(); }

size_t sb_max_compress_len(size_t input_len) {
    uint64_t n = (uint64_t)input_len;
    if (n > SB_MAX_INPUT) return 0;
    uint64_t m = 32 + n + n / 6;
    return m > SB_MAX_INPUT ? 0 : (size_t)m;
}

size_t sb_frame_max_len(size_t n) {
    size_t chunks = (n + SB_MAX_BLOCK - 1) / SB_MAX_BLOCK;
    return 10 + chunks * (8 + (size_t)SB_MAX_CBLOCK);
}

int sb_compress(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_n, sb_error* err) {
    if ((!in && n) || !out || !out_n) return fail(err, SB_E_INVALID);
    const size_t need = sb_max_compress_len(n);
    if (need == 0) return fail(err, SB_TOO_BIG, (uint64_t)n, SB_MAX_INPUT);
    if (cap < need) return fail(err, SB_BUFFER_TOO_SMALL, (uint64_t)cap, (uint64_t)need);
    Ctx* c;
    int rc = get_ctx(&c, err);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(c->in[0].need(n + 64));
    CK(c->compact[0].need(need + 64));
    if (n) CK(cudaMemcpyAsync(c->in[0].p, in, n, cudaMemcpyHostToDevice, c->s_compute));
    uint64_t total = 0;
    rc = compress_stream_device(*c, c->in[0].as<uint8_t>(), n, c->compact[0].as<uint8_t>(), 0, 0, &total, c->s_compute, err);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, c->compact[0].p, total, cudaMemcpyDeviceToHost, c->s_compute));
    CK(cudaStreamSynchronize(c->s_compute));
    *out_n = (size_t)total;
    ok(err);
    return 0;
}

int sb_decompress_len(const uint8_t* in, size_t n, size_t* out_len, sb_error* err) {
    // reference src/decompress.rs:30-35, 362-374 -- header arithmetic only
    if (!out_len || (!in && n)) return fail(err, SB_E_INVALID);
    if (n == 0) { *out_len = 0; ok(err); return 0; }
    uint64_t v;
    size_t h = get_varint(in, n, &v);
    if (h == 0) return fail(err, SB_HEADER);
    if (v > SB_MAX_INPUT) return fail(err, SB_TOO_BIG, v, SB_MAX_INPUT);
    *out_len = (size_t)v;
    ok(err);
    return 0;
}

int sb_decompress(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_n, sb_error* err) {
    if ((!in && n) || (!out && cap) || !out_n) return fail(err, SB_E_INVALID);
    if (n == 0) return fail(err, SB_EMPTY);
    if (n > SB_MAX_INPUT) return fail(err, SB_E_INVALID);
    // the header decides how much device output we need; the kernel re-validates everything
    uint64_t v = 0;
    size_t h = get_varint(in, n, &v);
    uint64_t dcap = (h && v <= SB_MAX_INPUT && v <= cap) ? v : 0;
    Ctx* c;
    int rc = get_ctx(&c, err);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(c->in[0].need(n + 64));
    CK(c->compact[0].need(dcap + 64));
    CK(c->status[0].need(sizeof(sb_error) + 16));
    CK(cudaMemcpyAsync(c->in[0].p, in, n, cudaMemcpyHostToDevice, c->s_compute));
    sb_batch b;
    memset(&b, 0, sizeof b);
    b.in_base = c->in[0].as<uint8_t>(); b.in_len_uniform = (uint32_t)n;
    b.out_base = c->compact[0].as<uint8_t>();
    b.out_cap_uniform = cap > SB_MAX_INPUT ? (uint32_t)SB_MAX_INPUT : (uint32_t)cap;
    b.statuses = c->status[0].as<sb_error>();
    b.out_lens = (uint32_t*)((uint8_t*)c->status[0].p + sizeof(sb_error));
    b.count = 1;
    rc = launch_k2(*c, b, c->s_compute, err);
    if (rc) return rc;
    struct { sb_error e; uint32_t len; uint32_t pad; } res;
    CK(cudaMemcpyAsync(&res, c->status[0].p, sizeof(sb_error) + 8, cudaMemcpyDeviceToHost, c->s_compute));
    CK(cudaStreamSynchronize(c->s_compute));
    if (res.e.code) { if (err) *err = res.e; return (int)res.e.code; }
    if (res.len) CK(cudaMemcpy(out, c->compact[0].p, res.len, cudaMemcpyDeviceToHost));
    *out_n = res.len;
    ok(err);
    return 0;
}

int sb_crc32c_masked(const uint8_t* in, size_t n, uint32_t* out, sb_error* err) {
    if ((!in && n) || !out || n > SB_MAX_INPUT) return fail(err, SB_E_INVALID);
    Ctx* c;
    int rc = get_ctx(&c, err);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(c->in[0].need(n + 64));
    CK(c->crcs[0].need(16));
    if (n) CK(cudaMemcpyAsync(c->in[0].p, in, n, cudaMemcpyHostToDevice, c->s_compute));
    sb_batch b;
    memset(&b, 0, sizeof b);
    b.in_base = c->in[0].as<uint8_t>(); b.in_len_uniform = (uint32_t)n;
    b.out_lens = c->crcs[0].as<uint32_t>(); b.count = 1;
    rc = launch_k3(*c, b, c->s_compute, err);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, c->crcs[0].p, 4, cudaMemcpyDeviceToHost, c->s_compute));
    CK(cudaStreamSynchronize(c->s_compute));
    ok(err);
    return 0;
}

// ---------------------------------------------------------- device batches
int sb_compress_batch_device(const sb_batch* batch, void* stream, sb_error* err) {
    if (!batch || !batch->out_lens) return fail(err, SB_E_INVALID);
    Ctx* c;
    int rc = get_ctx(&c, err);
    if (rc) return rc;
    rc = launch_k1(*c, *batch, 1u, (cudaStream_t)stream, err);
    if (rc) return rc;
    ok(err);
    return 0;
}

int sb_decompress_batch_device(const sb_batch* batch, void* stream, sb_error* err) {
    if (!batch) return fail(err, SB_E_INVALID);
    Ctx* c;
    int rc = get_ctx(&c, err);
    if (rc) return rc;
    rc = launch_k2(*c, *batch, (cudaStream_t)stream, err);
    if (rc) return rc;
    ok(err);
    return 0;
}

int sb_crc32c_masked_batch_device(const sb_batch* batch, void* stream, sb_error* err) {
    if (!batch || !batch->out_lens) return fail(err, SB_E_INVALID);
    Ctx* c;
    int rc = get_ctx(&c, err);
    if (rc) return rc;
    rc = launch_k3(*c, *batch, (cudaStream_t)stream, err);
    if (rc) return rc;
    ok(err);
    return 0;
}

int sb_generate_blocks_device(const uint8_t* d_text, uint64_t text_len, uint8_t* d_out, uint64_t stride,
                              uint32_t len, uint64_t first, uint64_t count, uint64_t mul, void* stream, sb_error* err) {
    if (!d_text || !d_out || text_len < len) return fail(err, SB_E_INVALID);
    Ctx* c;
    int rc = get_ctx(&c, err);
    if (rc) return rc;
    if (count == 0) return 0;
    sbk::GenPlan g{d_text, text_len, d_out, stride, len, first, count, mul};
    uint64_t blocks = (count + 7) / 8;
    unsigned grid = (unsigned)(16 * c->sms);
    if (grid > blocks) grid = (unsigned)blocks;
    k6_generate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g);
    g_launches++;
    CK(cudaGetLastError());
    ok(err);
    return 0;
}

// ------------------------------------------------------------ host batches
// Waves of units are staged H2D on one stream, run on a second, and drained D2H
// on a third, double buffered, so PCIe traffic overlaps the kernels.
namespace {
const size_t WAVE_BYTES = (size_t)1 << 30;

struct Wave { size_t first, count; uint64_t in_bytes; };

std::vector<Wave> plan_waves(const uint32_t* in_lens, size_t count, const uint32_t* out_caps) {
    std::vector<Wave> w;
    size_t i = 0;
    while (i < count) {
        Wave cur{i, 0, 0};
        uint64_t outb = 0;
        // ramp-up: the first waves are small so the first kernel starts after ~1 ms of H2D, not ~10
        const size_t limit = w.size() == 0 ? WAVE_BYTES / 16 : w.size() == 1 ? WAVE_BYTES / 4 : WAVE_BYTES;
        while (i < count && cur.count < (1u << 20)) {
            uint64_t add = in_lens[i], oadd = out_caps ? out_caps[i] : 0;
            if (cur.count && (cur.in_bytes + add > limit || outb + oadd > 2 * limit)) break;
            cur.in_bytes += add + 16; outb += oadd; cur.count++; i++;
        }
        w.push_back(cur);
    }
    return w;
}
}  // namespace

int sb_compress_batch_host(const uint8_t* in_base, const uint64_t* in_offs, const uint32_t* in_lens,
                           uint8_t* out_base, const uint64_t* out_offs, const uint32_t* out_caps,
                           uint32_t* out_lens, size_t count, sb_error* err) {
    if (!in_base || !in_offs || !in_lens || !out_base || !out_offs || !out_lens) return fail(err, SB_E_INVALID);
    for (size_t i = 0; i < count; i++) {
        if (in_lens[i] > SB_MAX_BLOCK) return fail(err, SB_E_INVALID, i);   // one block per unit in the batched form
        if (out_caps && out_caps[i] < sb_max_compress_len(in_lens[i]))
            return fail(err, SB_BUFFER_TOO_SMALL, out_caps[i], sb_max_compress_len(in_lens[i]));
    }
    Ctx* c;
    int rc = get_ctx(&c, err);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    std::vector<Wave> waves = plan_waves(in_lens, count, nullptr);
    cudaEvent_t ev_in[2], ev_k[2], ev_out[2];
    for (int k = 0; k < 2; k++) {
        CK(cudaEventCreateWithFlags(&ev_in[k], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ev_k[k], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ev_out[k], cudaEventDisableTiming));
    }
    std::vector<uint64_t> dev_in_off[2];
    auto stage_in = [&](size_t wi) -> int {
        const Wave& w = waves[wi];
        const int b = (int)(wi & 1);
        CK(c->in[b].need(w.in_bytes + 64));
        CK(c->slots[b].need(w.count * (size_t)sbk::kSlotStride));
        CK(c->lens[b].need(w.count * 4 + 4));
        CK(c->caps[b].need(w.count * 4 + 4));
        CK(c->ptrs_in[b].need(w.count * 8 + 8));
        // coalesce units that are contiguous on the host into single copies
        std::vector<uint64_t>& doff = dev_in_off[b];
        doff.resize(w.count);
        uint64_t at = 0;
        size_t i = 0;
        while (i < w.count) {
            size_t j = i;
            uint64_t run = 0;
            const uint64_t h0 = in_offs[w.first + i];
            while (j < w.count && in_offs[w.first + j] == h0 + run) { doff[j] = at + run; run += in_lens[w.first + j]; j++; }
            if (run) CK(cudaMemcpyAsync(c->in[b].as<uint8_t>() + at, in_base + h0, run, cudaMemcpyHostToDevice, c->s_h2d));
            at += (run + 15) & ~(uint64_t)15;
            i = j;
        }
        { int prc = need_pinned(*c, b, w.count * 12 + 64, err); if (prc) return prc; }
        uint64_t* ptrs = (uint64_t*)c->pinned[b];
        uint32_t* plen = (uint32_t*)(ptrs + w.count);
        for (size_t k = 0; k < w.count; k++) ptrs[k] = (uint64_t)(uintptr_t)(c->in[b].as<uint8_t>() + doff[k]);
        memcpy(plen, in_lens + w.first, w.count * 4);
        CK(cudaMemcpyAsync(c->ptrs_in[b].p, ptrs, w.count * 8, cudaMemcpyHostToDevice, c->s_h2d));
        CK(cudaMemcpyAsync(c->caps[b].p, plen, w.count * 4, cudaMemcpyHostToDevice, c->s_h2d));
        CK(cudaStreamSynchronize(c->s_h2d));   // host temporaries + simple ordering; copies of the NEXT wave overlap kernels
        CK(cudaEventRecord(ev_in[b], c->s_h2d));
        return 0;
    };
    const bool timing = getenv("SNAPB200_TIMING") != nullptr;
    auto now_ms = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t_begin = now_ms();
    if (!waves.empty()) { rc = stage_in(0); if (rc) return rc; }
    for (size_t wi = 0; wi < waves.size(); wi++) {
        const Wave& w = waves[wi];
        const int b = (int)(wi & 1);
        if (timing) fprintf(stderr, "[compress wave %zu] t=%.2f launch (count %zu)\n", wi, now_ms() - t_begin, w.count);
        CK(cudaStreamWaitEvent(c->s_compute, ev_in[b], 0));
        if (wi >= 2) CK(cudaStreamWaitEvent(c->s_compute, ev_out[b], 0));   // wave wi-2 (same buffers) fully drained
        sb_batch bt;
        memset(&bt, 0, sizeof bt);
        bt.in_ptrs = (const uint8_t* const*)c->ptrs_in[b].p; bt.in_lens = c->caps[b].as<uint32_t>();
        bt.out_base = c->slots[b].as<uint8_t>(); bt.out_stride = sbk::kSlotStride; bt.out_cap_uniform = sbk::kSlotStride;
        bt.out_lens = c->lens[b].as<uint32_t>(); bt.count = (uint32_t)w.count;
        cudaEvent_t tk0 = nullptr, tk1 = nullptr;
        if (timing) { cudaEventCreate(&tk0); cudaEventCreate(&tk1); cudaEventRecord(tk0, c->s_compute); }
        rc = launch_k1(*c, bt, 1u, c->s_compute, err);
        if (rc) return rc;
        if (timing) cudaEventRecord(tk1, c->s_compute);
        // results come back through pinned staging: a D2H copy into the caller's (pageable) array
        // would block this thread until the kernel is done and serialise the next wave's H2D behind it
        { int prc = need_pinned(*c, 2 + b, w.count * 4 + 64, err); if (prc) return prc; }
        CK(cudaMemcpyAsync(c->pinned[2 + b], c->lens[b].p, w.count * 4, cudaMemcpyDeviceToHost, c->s_compute));
        CK(cudaEventRecord(ev_k[b], c->s_compute));
        if (wi + 1 < waves.size()) { rc = stage_in(wi + 1); if (rc) return rc; }   // overlaps the kernel above
        if (timing) fprintf(stderr, "[compress wave %zu] t=%.2f staged next\n", wi, now_ms() - t_begin);
        CK(cudaEventSynchronize(ev_k[b]));
        memcpy(out_lens + w.first, c->pinned[2 + b], w.count * 4);
        if (timing) { float kms = 0; cudaEventElapsedTime(&kms, tk0, tk1); fprintf(stderr, "[compress wave %zu] t=%.2f kernel done (kernel %.2f ms)\n", wi, now_ms() - t_begin, kms); cudaEventDestroy(tk0); cudaEventDestroy(tk1); }
        // drain: contiguous host destinations are gathered on the device first, then one D2H
        bool dense = true;
        uint64_t run = 0;
        for (size_t k = 0; k < w.count && dense; k++) {
            if (out_offs[w.first + k] != out_offs[w.first] + run) dense = false;
            run += out_lens[w.first + k];
        }
        if (dense && w.count > 1) {
            CK(c->compact[b].need(run + 64));
            CK(c->csize[b].need(w.count * 4 + 4));
            CK(c->offs[b].need((w.count + 1) * 8));
            sbk::FramePlan p;
            memset(&p, 0, sizeof p);
            p.slots = c->slots[b].as<uint8_t>(); p.clens = c->lens[b].as<uint32_t>(); p.nchunks = (uint32_t)w.count;
            p.frame = 0; p.base = 0; p.csize = c->csize[b].as<uint32_t>(); p.offs = c->offs[b].as<uint64_t>();
            p.out = c->compact[b].as<uint8_t>();
            rc = launch_assemble(*c, p, c->s_compute, err);
            if (rc) return rc;
            CK(cudaEventRecord(ev_k[b], c->s_compute));
            CK(cudaStreamWaitEvent(c->s_d2h, ev_k[b], 0));
            CK(cudaMemcpyAsync(out_base + out_offs[w.first], c->compact[b].p, run, cudaMemcpyDeviceToHost, c->s_d2h));
        } else {
            for (size_t k = 0; k < w.count; k++)
                CK(cudaMemcpyAsync(out_base + out_offs[w.first + k], c->slots[b].as<uint8_t>() + k * (size_t)sbk::kSlotStride,
                                   out_lens[w.first + k], cudaMemcpyDeviceToHost, c->s_d2h));
        }
        CK(cudaEventRecord(ev_out[b], c->s_d2h));
    }
    CK(cudaStreamSynchronize(c->s_d2h));
    CK(cudaStreamSynchronize(c->s_compute));
    for (int k = 0; k < 2; k++) { cudaEventDestroy(ev_in[k]); cudaEventDestroy(ev_k[k]); cudaEventDestroy(ev_out[k]); }
    ok(err);
    return 0;
}

int sb_decompress_batch_host(const uint8_t* in_base, const uint64_t* in_offs, const uint32_t* in_lens,
                             uint8_t* out_base, const uint64_t* out_offs, const uint32_t* out_caps,
                             uint32_t* out_lens, sb_error* statuses, size_t count, sb_error* err) {
    if (!in_base || !in_offs || !in_lens || !out_base || !out_offs || !out_caps || !out_lens || !statuses)
        return fail(err, SB_E_INVALID);
    Ctx* c;
    int rc = get_ctx(&c, err);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    std::vector<Wave> waves = plan_waves(in_lens, count, out_caps);
    cudaEvent_t ev_in[2], ev_k[2], ev_out[2];
    for (int k = 0; k < 2; k++) {
        CK(cudaEventCreateWithFlags(&ev_in[k], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ev_k[k], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ev_out[k], cudaEventDisableTiming));
    }
    std::vector<uint64_t> pin[2], pout[2];
    // H2D of wave wi into buffer set wi&1 (copy stream; overlaps the previous wave's kernel)
    auto stage_in = [&](size_t wi) -> int {
        const Wave& w = waves[wi];
        const int b = (int)(wi & 1);
        uint64_t out_total = 0;
        for (size_t k = 0; k < w.count; k++) out_total += ((uint64_t)out_caps[w.first + k] + 15) & ~(uint64_t)15;
        CK(c->in[b].need(w.in_bytes + 64));
        CK(c->compact[b].need(out_total + 64));
        CK(c->lens[b].need(w.count * 4 + 4));
        CK(c->caps[b].need(w.count * 8 + 8));
        CK(c->status[b].need(w.count * sizeof(sb_error)));
        CK(c->ptrs_in[b].need(w.count * 8 + 8));
        CK(c->ptrs_out[b].need(w.count * 8 + 8));
        pin[b].resize(w.count); pout[b].resize(w.count);
        uint64_t at = 0, oat = 0;
        size_t i = 0;
        while (i < w.count) {                                      // host-contiguous units travel as one copy
            size_t j = i;
            uint64_t run = 0;
            const uint64_t h0 = in_offs[w.first + i];
            while (j < w.count && in_offs[w.first + j] == h0 + run) {
                pin[b][j] = (uint64_t)(uintptr_t)(c->in[b].as<uint8_t>() + at + run); run += in_lens[w.first + j]; j++;
            }
            if (run) CK(cudaMemcpyAsync(c->in[b].as<uint8_t>() + at, in_base + h0, run, cudaMemcpyHostToDevice, c->s_h2d));
            at += (run + 15) & ~(uint64_t)15;
            i = j;
        }
        for (size_t k = 0; k < w.count; k++) {
            pout[b][k] = (uint64_t)(uintptr_t)(c->compact[b].as<uint8_t>() + oat);
            oat += ((uint64_t)out_caps[w.first + k] + 15) & ~(uint64_t)15;
        }
        { int prc = need_pinned(*c, b, w.count * 24 + 64, err); if (prc) return prc; }
        uint64_t* sp = (uint64_t*)c->pinned[b];
        memcpy(sp, pin[b].data(), w.count * 8);
        memcpy(sp + w.count, pout[b].data(), w.count * 8);
        uint32_t* sl = (uint32_t*)(sp + 2 * w.count);
        memcpy(sl, in_lens + w.first, w.count * 4);
        memcpy(sl + w.count, out_caps + w.first, w.count * 4);
        CK(cudaMemcpyAsync(c->ptrs_in[b].p, sp, w.count * 8, cudaMemcpyHostToDevice, c->s_h2d));
        CK(cudaMemcpyAsync(c->ptrs_out[b].p, sp + w.count, w.count * 8, cudaMemcpyHostToDevice, c->s_h2d));
        CK(cudaMemcpyAsync(c->caps[b].p, sl, w.count * 8, cudaMemcpyHostToDevice, c->s_h2d));
        CK(cudaStreamSynchronize(c->s_h2d));
        CK(cudaEventRecord(ev_in[b], c->s_h2d));
        return 0;
    };
    if (!waves.empty()) { rc = stage_in(0); if (rc) return rc; }
    for (size_t wi = 0; wi < waves.size(); wi++) {
        const Wave& w = waves[wi];
        const int b = (int)(wi & 1);
        CK(cudaStreamWaitEvent(c->s_compute, ev_in[b], 0));
        if (wi >= 2) CK(cudaStreamWaitEvent(c->s_compute, ev_out[b], 0));   // wave wi-2 (same buffers) fully drained
        sb_batch bt;
        memset(&bt, 0, sizeof bt);
        bt.in_ptrs = (const uint8_t* const*)c->ptrs_in[b].p; bt.in_lens = c->caps[b].as<uint32_t>();
        bt.out_ptrs = (uint8_t* const*)c->ptrs_out[b].p; bt.out_caps = c->caps[b].as<uint32_t>() + w.count;
        bt.out_lens = c->lens[b].as<uint32_t>(); bt.statuses = c->status[b].as<sb_error>(); bt.count = (uint32_t)w.count;
        rc = launch_k2(*c, bt, c->s_compute, err);
        if (rc) return rc;
        { int prc = need_pinned(*c, 2 + b, w.count * (4 + sizeof(sb_error)) + 64, err); if (prc) return prc; }
        sb_error* pst = (sb_error*)c->pinned[2 + b];
        uint32_t* pln = (uint32_t*)(pst + w.count);
        CK(cudaMemcpyAsync(pln, c->lens[b].p, w.count * 4, cudaMemcpyDeviceToHost, c->s_compute));
        CK(cudaMemcpyAsync(pst, c->status[b].p, w.count * sizeof(sb_error), cudaMemcpyDeviceToHost, c->s_compute));
        CK(cudaEventRecord(ev_k[b], c->s_compute));
        if (wi + 1 < waves.size()) { rc = stage_in(wi + 1); if (rc) return rc; }   // overlaps the kernel above and the previous drain
        CK(cudaEventSynchronize(ev_k[b]));
        memcpy(out_lens + w.first, pln, w.count * 4);
        memcpy(statuses + w.first, pst, w.count * sizeof(sb_error));
        // drain on the third stream; contiguous destinations whose caps are exactly filled go out as one copy
        const uint64_t cbase = (uint64_t)(uintptr_t)c->compact[b].p;
        size_t k = 0;
        while (k < w.count) {
            size_t j = k;
            uint64_t run = 0;
            const uint64_t h0 = out_offs[w.first + k];
            const uint64_t d0 = pout[b][k] - cbase;
            while (j < w.count && out_offs[w.first + j] == h0 + run && pout[b][j] - cbase == d0 + run) {
                run += out_lens[w.first + j];
                const bool full = out_lens[w.first + j] == out_caps[w.first + j] && (out_caps[w.first + j] & 15u) == 0;
                j++;
                if (!full) break;
            }
            if (run) CK(cudaMemcpyAsync(out_base + h0, c->compact[b].as<uint8_t>() + d0, run, cudaMemcpyDeviceToHost, c->s_d2h));
            k = j;
        }
        CK(cudaEventRecord(ev_out[b], c->s_d2h));
    }
    CK(cudaStreamSynchronize(c->s_d2h));
    CK(cudaStreamSynchronize(c->s_compute));
    for (int k = 0; k < 2; k++) { cudaEventDestroy(ev_in[k]); cudaEventDestroy(ev_k[k]); cudaEventDestroy(ev_out[k]); }
    ok(err);
    return 0;
}

// -------------------------------------------------------------- frame format
int sb_frame_encode_device(const uint8_t* d_in, uint64_t n, uint8_t* d_out, uint64_t cap,
                           int include_ident, uint64_t* out_n, void* stream, sb_error* err) {
    if ((!d_in && n) || !out_n || (!d_out && n)) return fail(err, SB_E_INVALID);
    if (cap < sb_frame_max_len(n) - (include_ident ? 0 : 10)) return fail(err, SB_BUFFER_TOO_SMALL, cap, sb_frame_max_len(n));
    Ctx* c;
    int rc = get_ctx(&c, err);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    rc = compress_stream_device(*c, d_in, n, d_out, 1, include_ident, out_n, stream ? (cudaStream_t)stream : c->s_compute, err);
    if (rc) return rc;
    ok(err);
    return 0;
}

int sb_frame_encode_ex(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_n, int include_ident, sb_error* err);
int sb_frame_encode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_n, sb_error* err) {
    return sb_frame_encode_ex(in, n, out, cap, out_n, 1, err);
}
int sb_frame_encode_ex(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_n, int include_ident, sb_error* err) {
    if ((!in && n) || !out_n || (!out && n)) return fail(err, SB_E_INVALID);
    if (n == 0) { *out_n = 0; ok(err); return 0; }              // src/write.rs:155-157: nothing is written
    if (cap < sb_frame_max_len(n)) return fail(err, SB_BUFFER_TOO_SMALL, cap, sb_frame_max_len(n));
    Ctx* c;
    int rc = get_ctx(&c, err);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(c->in[1].need(n + 64));
    CK(c->compact[1].need(sb_frame_max_len(n) + 64));
    CK(cudaMemcpyAsync(c->in[1].p, in, n, cudaMemcpyHostToDevice, c->s_compute));
    uint64_t total = 0;
    rc = compress_stream_device(*c, c->in[1].as<uint8_t>(), n, c->compact[1].as<uint8_t>(), 1, include_ident, &total, c->s_compute, err);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, c->compact[1].p, total, cudaMemcpyDeviceToHost, c->s_compute));
    CK(cudaStreamSynchronize(c->s_compute));
    *out_n = (size_t)total;
    ok(err);
    return 0;
}

// read::FrameDecoder + read_to_end over host memory (reference src/read.rs:104-239).
// The host walks the chunk headers (each one gives the next offset), the device
// decodes every compressed chunk (K2), copies uncompressed ones, checksums all
// outputs (K3); the first failure IN STREAM ORDER is reported.
int sb_frame_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_n, sb_error* err) {
    if ((!in && n) || !out_n) return fail(err, SB_E_INVALID);
    struct Chunk { uint64_t body_off; uint32_t body_len; uint32_t dlen; uint32_t want_crc; uint8_t type; };
    std::vector<Chunk> chunks;
    // shadow of the decoder's persistent 76490-byte `src` buffer, needed only for the
    // reference quirk that decompress_len() is applied to the WHOLE buffer (src/read.rs:216)
    std::vector<uint8_t> shadow(SB_MAX_CBLOCK, 0);
    sb_error walk_err;
    memset(&walk_err, 0, sizeof walk_err);
    size_t pos = 0;
    uint64_t produced = 0;
    bool seen_ident = false;
    auto werr = [&](uint32_t code, uint64_t a = 0, uint64_t b = 0) { walk_err.code = code; walk_err.a = a; walk_err.b = b; };
    while (pos < n) {
        if (n - pos < 4) { werr(SB_IO_UNEXPECTED_EOF); break; }
        const uint8_t* h = in + pos;
        memcpy(shadow.data(), h, 4);
        pos += 4;
        const uint8_t ty = h[0];
        if (!seen_ident) {
            if (ty != 0xFF) { werr(SB_STREAM_HEADER, ty); break; }
            seen_ident = true;
        }
        const uint64_t len = (uint64_t)h[1] | ((uint64_t)h[2] << 8) | ((uint64_t)h[3] << 16);
        if (len > SB_MAX_CBLOCK) { werr(SB_UNSUPPORTED_CHUNK_LENGTH, len, 0); break; }
        if (ty >= 0x02 && ty <= 0x7F) { werr(SB_UNSUPPORTED_CHUNK_TYPE, ty); break; }
        if ((ty >= 0x80 && ty <= 0xFD) || ty == 0xFE) {
            if (n - pos < len) { werr(SB_IO_UNEXPECTED_EOF); break; }
            memcpy(shadow.data(), in + pos, len);
            pos += len;
        } else if (ty == 0xFF) {
            if (len != 6) { werr(SB_UNSUPPORTED_CHUNK_LENGTH, len, 1); break; }
            if (n - pos < 6) { werr(SB_IO_UNEXPECTED_EOF); break; }
            memcpy(shadow.data(), in + pos, 6);
            if (memcmp(in + pos, "sNaPpY", 6) != 0) {
                uint64_t a = 0;
                for (int i = 0; i < 6; i++) a |= (uint64_t)in[pos + i] << (8 * i);
                werr(SB_STREAM_HEADER_MISMATCH, a);
                break;
            }
            pos += 6;
        } else {
            if (len < 4) { werr(SB_UNSUPPORTED_CHUNK_LENGTH, len, 0); break; }
            if (n - pos < 4) { werr(SB_IO_UNEXPECTED_EOF); break; }
            uint32_t want;
            memcpy(&want, in + pos, 4);
            pos += 4;
            const uint32_t body = (uint32_t)len - 4;
            Chunk ch{pos, body, 0, want, ty};
            if (ty == 0x01) {
                if (body > SB_MAX_BLOCK) { werr(SB_UNSUPPORTED_CHUNK_LENGTH, body, 0); break; }
                if (n - pos < body) { werr(SB_IO_UNEXPECTED_EOF); break; }
                ch.dlen = body;
            } else {
                if (n - pos < body) { werr(SB_IO_UNEXPECTED_EOF); break; }
                // decompress_len over the persistent buffer: only the first <=10 bytes matter
                uint8_t head[16];
                const size_t fresh = body < 16 ? body : 16;
                memcpy(head, in + pos, fresh);
                if (fresh < 16) memcpy(head + fresh, shadow.data() + fresh, 16 - fresh);
                uint64_t v = 0;
                const size_t hl = get_varint(head, 16, &v);   // a varint never needs more than 10 bytes
                if (hl == 0) { werr(SB_HEADER); break; }
                if (v > SB_MAX_INPUT) { werr(SB_TOO_BIG, v, SB_MAX_INPUT); break; }
                if (v > SB_MAX_BLOCK) { werr(SB_UNSUPPORTED_CHUNK_LENGTH, v, 0); break; }
                ch.dlen = (uint32_t)v;
                const size_t keep = body < 16 ? body : 16;   // later quirk reads only look at the first bytes
                memcpy(shadow.data(), in + pos, keep);
            }
            pos += body;
            chunks.push_back(ch);
            produced += ch.dlen;
        }
    }
    // Sizing call: only possible failures that precede any data check are reported by the full call.
    if (!out) { *out_n = (size_t)produced; ok(err); return 0; }
    if (produced > cap) return fail(err, SB_BUFFER_TOO_SMALL, cap, produced);

    uint64_t good = 0;      // bytes produced by chunks before the first failing chunk
    sb_error first;
    memset(&first, 0, sizeof first);
    if (!chunks.empty()) {
        Ctx* c;
        int rc = get_ctx(&c, err);
        if (rc) return rc;
        std::lock_guard<std::mutex> lk(c->mu);
        const size_t m = chunks.size();
        CK(c->in[1].need(n + 64));
        CK(c->compact[1].need(produced + 64));
        CK(c->ptrs_in[1].need(m * 8 + 8));
        CK(c->ptrs_out[1].need(m * 8 + 8));
        CK(c->caps[1].need(m * 8 + 8));
        CK(c->lens[1].need(m * 4 + 4));
        CK(c->crcs[1].need(m * 4 + 4));
        CK(c->status[1].need(m * sizeof(sb_error)));
        CK(cudaMemcpyAsync(c->in[1].p, in, n, cudaMemcpyHostToDevice, c->s_compute));
        // compressed chunks first, then uncompressed ones (two sub-batches sharing the arrays)
        std::vector<uint64_t> pin(m), pout(m);
        std::vector<uint32_t> ilen(m), ocap(m);
        std::vector<size_t> order;
        order.reserve(m);
        for (size_t i = 0; i < m; i++) if (chunks[i].type == 0x00) order.push_back(i);
        const size_t ncomp = order.size();
        for (size_t i = 0; i < m; i++) if (chunks[i].type == 0x01) order.push_back(i);
        std::vector<uint64_t> ooff(m);
        uint64_t at = 0;
        for (size_t i = 0; i < m; i++) { ooff[i] = at; at += chunks[i].dlen; }
        for (size_t k = 0; k < m; k++) {
            const Chunk& ch = chunks[order[k]];
            pin[k] = (uint64_t)(uintptr_t)(c->in[1].as<uint8_t>() + ch.body_off);
            pout[k] = (uint64_t)(uintptr_t)(c->compact[1].as<uint8_t>() + ooff[order[k]]);
            ilen[k] = ch.body_len;
            ocap[k] = ch.dlen;
        }
        CK(cudaMemcpyAsync(c->ptrs_in[1].p, pin.data(), m * 8, cudaMemcpyHostToDevice, c->s_compute));
        CK(cudaMemcpyAsync(c->ptrs_out[1].p, pout.data(), m * 8, cudaMemcpyHostToDevice, c->s_compute));
        CK(cudaMemcpyAsync(c->caps[1].p, ilen.data(), m * 4, cudaMemcpyHostToDevice, c->s_compute));
        CK(cudaMemcpyAsync(c->caps[1].as<uint32_t>() + m, ocap.data(), m * 4, cudaMemcpyHostToDevice, c->s_compute));
        CK(cudaMemsetAsync(c->status[1].p, 0, m * sizeof(sb_error), c->s_compute));
        sb_batch bt;
        memset(&bt, 0, sizeof bt);
        bt.in_ptrs = (const uint8_t* const*)c->ptrs_in[1].p; bt.in_lens = c->caps[1].as<uint32_t>();
        bt.out_ptrs = (uint8_t* const*)c->ptrs_out[1].p; bt.out_caps = c->caps[1].as<uint32_t>() + m;
        bt.out_lens = c->lens[1].as<uint32_t>(); bt.statuses = c->status[1].as<sb_error>(); bt.count = (uint32_t)ncomp;
        rc = launch_k2(*c, bt, c->s_compute, err);
        if (rc) return rc;
        sb_batch cp = bt;
        cp.in_ptrs += ncomp; cp.in_lens += ncomp; cp.out_ptrs += ncomp; cp.out_caps += ncomp;
        cp.out_lens = nullptr; cp.statuses = nullptr; cp.count = (uint32_t)(m - ncomp);
        rc = launch_copy_units(*c, cp, c->s_compute, err);
        if (rc) return rc;
        // checksum every produced chunk: unit k = output of order[k], length dlen
        sb_batch cr;
        memset(&cr, 0, sizeof cr);
        cr.in_ptrs = (const uint8_t* const*)c->ptrs_out[1].p; cr.in_lens = c->caps[1].as<uint32_t>() + m;
        cr.out_lens = c->crcs[1].as<uint32_t>(); cr.count = (uint32_t)m;
        rc = launch_k3(*c, cr, c->s_compute, err);
        if (rc) return rc;
        std::vector<sb_error> st(m);
        std::vector<uint32_t> crc(m);
        CK(cudaMemcpyAsync(st.data(), c->status[1].p, m * sizeof(sb_error), cudaMemcpyDeviceToHost, c->s_compute));
        CK(cudaMemcpyAsync(crc.data(), c->crcs[1].p, m * 4, cudaMemcpyDeviceToHost, c->s_compute));
        CK(cudaStreamSynchronize(c->s_compute));
        std::vector<sb_error> by_chunk(m);
        for (size_t k = 0; k < m; k++) {
            sb_error e = st[k];
            if (k >= ncomp) memset(&e, 0, sizeof e);
            if (!e.code && crc[k] != chunks[order[k]].want_crc) { e.code = SB_CHECKSUM; e.a = chunks[order[k]].want_crc; e.b = crc[k]; }
            by_chunk[order[k]] = e;
        }
        good = produced;
        for (size_t i = 0; i < m; i++) if (by_chunk[i].code) { first = by_chunk[i]; good = ooff[i]; break; }
        if (good) CK(cudaMemcpy(out, c->compact[1].p, good, cudaMemcpyDeviceToHost));
    }
    *out_n = (size_t)good;
    if (first.code) { if (err) *err = first; return (int)first.code; }
    if (walk_err.code) { if (err) *err = walk_err; return (int)walk_err.code; }
    ok(err);
    return 0;
}

// ---------------------------------------------------------- libsnappy C API
int snappy_compress(const char* input, size_t input_length, char* compressed, size_t* compressed_length) {
    if (!compressed_length) return 1;
    sb_error e;
    size_t n = 0;
    int rc = sb_compress((const uint8_t*)input, input_length, (uint8_t*)compressed, *compressed_length, &n, &e);
    if (rc == SB_BUFFER_TOO_SMALL) return 2;
    if (rc) return 1;
    *compressed_length = n;
    return 0;
}
int snappy_uncompress(const char* compressed, size_t compressed_length, char* uncompressed, size_t* uncompressed_length) {
    if (!uncompressed_length) return 1;
    sb_error e;
    size_t n = 0;
    int rc = sb_decompress((const uint8_t*)compressed, compressed_length, (uint8_t*)uncompressed, *uncompressed_length, &n, &e);
    if (rc == SB_BUFFER_TOO_SMALL) return 2;
    if (rc) return 1;
    *uncompressed_length = n;
    return 0;
}
size_t snappy_max_compressed_length(size_t source_length) { return 32 + source_length + source_length / 6; }
int snappy_uncompressed_length(const char* compressed, size_t compressed_length, size_t* result) {
    sb_error e;
    if (!result || compressed_length == 0) return 1;
    return sb_decompress_len((const uint8_t*)compressed, compressed_length, result, &e) ? 1 : 0;
}

#ifdef K1_PROFILE
// profile build only (tools/k1_phase_profile.sh): read / reset the parser phase timers
int sb_debug_k1_profile(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    if (out16 && cudaMemcpyFromSymbol(out16, g_k1_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return 1;
    if (reset) { unsigned long long z[16] = {0}; if (cudaMemcpyToSymbol(g_k1_prof, z, sizeof z) != cudaSuccess) return 1; }
    return 0;
}
#endif

}  // extern "C"
