// emu_kernels.cpp -- TEST TOOLING ONLY. Compiles the CUDA kernel bodies of
// rust-snappy_b200/csrc with g++ against the fiber warp emulator and exposes
// them to pytest through a C interface (tests/test_emu_kernels.py).
#define SB_EMU 1
#include "simt_emu.h"
#include "../../rust-snappy_b200/csrc/k1_compress.cuh"
#include "../../rust-snappy_b200/csrc/k2_decompress.cuh"

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

int emu_decompress_batch(const sb_batch* b, unsigned grid, unsigned block) {
    sb_batch c = *b;
    sbemu::launch(grid, block, (block / 32) * sbk::K2_SMEM_PER_WARP, k2_entry, &c);
    return 0;
}

}
