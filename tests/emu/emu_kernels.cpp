// emu_kernels.cpp -- TEST TOOLING ONLY. Compiles the CUDA kernel bodies of
// rust-snappy_b200/csrc with g++ against the fiber warp emulator and exposes
// them to pytest through a C interface (tests/test_emu_kernels.py).
#define SB_EMU 1
#include "simt_emu.h"
#include "../../rust-snappy_b200/csrc/k1_compress.cuh"
#include "../../rust-snappy_b200/csrc/k2_decompress.cuh"
#include "../../rust-snappy_b200/csrc/k4_frame.cuh"
#include "../../rust-snappy_b200/csrc/k5_frame_decode.cuh"

struct K1Args { sb_batch b; uint32_t flags; uint64_t* rings; uint16_t* gtables; uint32_t* work; uint32_t* crcs; };
static void k1_entry(void* a) {
    K1Args* x = (K1Args*)a;
    // 0x400: 7 shared-memory-table chains + 4 chains with tables in global memory; otherwise 7 + 0
    if (x->flags & 0x400u) sbk::k1_compress_body_multi<7, 4>(x->b, x->flags & 0xFFu, x->rings, x->gtables, x->work, x->crcs);
    else sbk::k1_compress_body_multi<7, 0>(x->b, x->flags & 0xFFu, x->rings, x->gtables, x->work, x->crcs);
}
static void k2_entry(void* a) { sbk::k2_decompress_body(*(sb_batch*)a); }

extern "C" {

// flags: bit0 = varint header, 0x400 = hybrid layout, bits 20..23 = chains per CTA (0 = all)
int emu_compress_batch(const sb_batch* b, uint32_t flags, unsigned grid, uint32_t* crcs) {
    K1Args a{*b, flags, nullptr, nullptr, nullptr, crcs};
    const unsigned ng = (flags & 0x400u) ? 4 : 0;
    std::vector<uint64_t> rings((size_t)grid * (7 + ng) * sbk::K1_RING_GW, 0xCDCDCDCDCDCDCDCDull);
    std::vector<uint16_t> gt((size_t)grid * (ng + 1) * (sbk::K1_TABLE_BYTES / 2) + 8, 0xCDCD);
    uint32_t work = 0;
    a.rings = rings.data();
    a.gtables = (uint16_t*)(((uintptr_t)gt.data() + 15) & ~(uintptr_t)15);
    a.work = &work;
    unsigned chains = (flags >> 20) & 15u;
    if (chains == 0 || chains > 7 + ng) chains = 7 + ng;
    sbemu::launch(grid, chains * 64, sbk::k1_multi_smem(7, ng), k1_entry, &a);
    return 0;
}

// ---- frame path (K4 assembly around K1, K5 decode) with the same kernel sequence as csrc/snapb200.cu
static size_t up256(size_t v) { return (v + 255) / 256 * 256; }
static void k4_fill_entry(void* a) { auto* x = (std::pair<sbk::FramePlan, uint32_t*>*)a; sbk::k4_fill_lens_body(x->second, x->first.n, x->first.nchunks); }
static void k4_scan_local_entry(void* a) { sbk::k4_scan_local_body(*(sbk::FramePlan*)a); }
static void k4_scan_tiles_entry(void* a) { sbk::k4_scan_tiles_body(*(sbk::FramePlan*)a); }
static void k4_gather_entry(void* a) { sbk::k4_gather_body(*(sbk::FramePlan*)a); }
static void k5_parse_entry(void* a) { sbk::k5_parse_body(*(sbk::DecodePlan*)a); }
static void k5_walk_entry(void* a) { sbk::k5_walk_body(*(sbk::DecodePlan*)a); }
static void k5_scan_local_entry(void* a) { sbk::k5_scan_local_body(*(sbk::DecodePlan*)a); }
static void k5_scan_tiles_entry(void* a) { sbk::k5_scan_tiles_body(*(sbk::DecodePlan*)a); }
static void k5_decode_entry(void* a) { sbk::k5_decode_body(*(sbk::DecodePlan*)a); }
static void k5_finish_entry(void* a) { sbk::k5_finish_body(*(sbk::DecodePlan*)a); }

// sb_frame_encode_device_ws under the emulator: fill lens -> K1 (+ chunk CRC in the emitter) -> scan -> gather
int emu_frame_encode(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, int ident, uint64_t* offs_out, sb_frame_result* result) {
    const uint32_t nchunks = (uint32_t)((n + 65535) / 65536);
    memset(result, 0, sizeof *result);
    if (n == 0) return 0;
    std::vector<uint8_t> slots((size_t)nchunks * sbk::kSlotStride + 64, 0xEE);
    std::vector<uint32_t> lens_in(nchunks + 1), clens(nchunks + 1), crcs(nchunks + 1);
    std::vector<uint64_t> offs(nchunks + 2), tiles(nchunks / sbk::K4_TILE + 4);
    sbk::FramePlan p;
    memset(&p, 0, sizeof p);
    if (ident) { memcpy(p.head, "\xff\x06\x00\x00sNaPpY", 10); p.head_len = 10; }
    p.in = in; p.n = n; p.slots = slots.data(); p.clens = clens.data(); p.crcs = crcs.data(); p.nchunks = nchunks; p.frame = 1;
    p.offs = offs_out ? offs_out : offs.data(); p.tiles = tiles.data(); p.out = out; p.cap = cap; p.result = result;
    std::pair<sbk::FramePlan, uint32_t*> fl{p, lens_in.data()};
    sbemu::launch((nchunks + 255) / 256, 256, 0, k4_fill_entry, &fl);
    sb_batch b;
    memset(&b, 0, sizeof b);
    b.in_base = in; b.in_stride = 65536; b.in_lens = lens_in.data();
    b.out_base = slots.data(); b.out_stride = sbk::kSlotStride; b.out_cap_uniform = sbk::kSlotStride; b.out_lens = clens.data(); b.count = nchunks;
    emu_compress_batch(&b, 1u | 0x400u, 2, crcs.data());
    sbemu::launch((nchunks + sbk::K4_TILE - 1) / sbk::K4_TILE, sbk::K4_TILE, 128, k4_scan_local_entry, &p);
    sbemu::launch(1, 1024, 1024 * 8, k4_scan_tiles_entry, &p);
    sbemu::launch(2, 256, 0, k4_gather_entry, &p);
    return 0;
}

// sb_frame_decode_device_ws under the emulator: [parse] -> walk -> scan -> decode+CRC -> finish
int emu_frame_decode(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, const uint64_t* index, uint32_t index_n, int fragment,
                     sb_frame_result* result, uint32_t max_chunks) {
    std::vector<uint8_t> scratch(up256((size_t)max_chunks * sizeof(sbk::FChunk) + 64) + up256(((size_t)max_chunks + 1) * 8) +
                                 up256(((size_t)max_chunks / sbk::K4_TILE + 3) * 8) + up256((size_t)max_chunks * sizeof(sb_error) + 64) + 1024, 0xCD);
    uint8_t* q = (uint8_t*)up256((size_t)scratch.data());
    sbk::DecodePlan p;
    memset(&p, 0, sizeof p);
    p.in = in; p.n = n; p.index = index; p.index_n = index ? index_n : 0; p.fragment = fragment ? 1u : 0u;
    p.chunks = (sbk::FChunk*)q; q += up256((size_t)max_chunks * sizeof(sbk::FChunk) + 64);
    p.ooff = (uint64_t*)q; q += up256(((size_t)max_chunks + 1) * 8);
    p.tiles = (uint64_t*)q; q += up256(((size_t)max_chunks / sbk::K4_TILE + 3) * 8);
    p.statuses = (sb_error*)q; q += up256((size_t)max_chunks * sizeof(sb_error) + 64);
    p.ctl = (sbk::DecodeCtl*)q;
    p.cap_chunks = max_chunks; p.out = out; p.cap = cap; p.result = result;
    memset(p.ctl, 0, sizeof(sbk::DecodeCtl));
    if (p.index) sbemu::launch(p.index_n ? (p.index_n + 255) / 256 : 1, 256, 0, k5_parse_entry, &p);
    sbemu::launch(1, 32, 0, k5_walk_entry, &p);
    const unsigned ntiles = (max_chunks + sbk::K4_TILE - 1) / sbk::K4_TILE;
    sbemu::launch(ntiles ? ntiles : 1, sbk::K4_TILE, 128, k5_scan_local_entry, &p);
    sbemu::launch(1, 1024, 1024 * 8, k5_scan_tiles_entry, &p);
    sbemu::launch(3, 128, sbk::K3_TABLE_BYTES + 4 * sbk::K2_SMEM_PER_WARP, k5_decode_entry, &p);
    sbemu::launch(1, 32, 0, k5_finish_entry, &p);
    return 0;
}

int emu_decompress_batch(const sb_batch* b, unsigned grid, unsigned block) {
    sb_batch c = *b;
    sbemu::launch(grid, block, (block / 32) * sbk::K2_SMEM_PER_WARP, k2_entry, &c);
    return 0;
}

}
