// emu_kernels.cpp -- TEST TOOLING ONLY. Compiles the CUDA kernel bodies of
// rust-snappy_b200/csrc with g++ against the fiber warp emulator and exposes
// them to pytest through a C interface (tests/test_emu_kernels.py).
#define SB_EMU 1
#include "../../rust-snappy_b200/csrc/k1_compress.cuh"
#include "../../rust-snappy_b200/csrc/k2_decompress.cuh"

struct K1Args { sb_batch b; uint32_t flags; uint64_t* rings; uint16_t* gtables; uint32_t* work; };
static void k1_entry(void* a) {
    K1Args* x = (K1Args*)a;
    if (x->flags & 0x1000000u) { sbk::k1_compress_body_multi<6, 8, true>(x->b, x->flags & 0xFFu, x->rings, x->gtables, x->work); return; }   // second-generation parser: 6 smem + 8 global tables
    if (x->flags & 0x400u) { sbk::k1_compress_body_multi<7, 4>(x->b, (x->flags & 0xFFu) | ((x->flags & 0x800u) ? 8u : 0u), x->rings, x->gtables, x->work); return; }   // hybrid; 0x800 = mbarrier wake-up: 7 smem + 4 global tables
    if (x->flags & 0x200u) { sbk::k1_compress_body_multi<7, 0>(x->b, x->flags & 0xFFu, x->rings, x->gtables, x->work); return; }
    const bool gw = x->flags & 0x100u;
    const unsigned np = (x->flags >> 12) & 7u, f = x->flags & 0xFFu;
    if (gw) { if (np <= 1) sbk::k1_compress_body<true, 1>(x->b, f); else if (np == 2) sbk::k1_compress_body<true, 2>(x->b, f); else sbk::k1_compress_body<true, 3>(x->b, f); }
    else { if (np <= 1) sbk::k1_compress_body<false, 1>(x->b, f); else if (np == 2) sbk::k1_compress_body<false, 2>(x->b, f); else sbk::k1_compress_body<false, 3>(x->b, f); }
}
static void k2_entry(void* a) { sbk::k2_decompress_body(*(sb_batch*)a); }

extern "C" {

int emu_compress_batch(const sb_batch* b, uint32_t flags, unsigned grid) {
    K1Args a{*b, flags, nullptr, nullptr, nullptr};
    sbk::g_k1_gt_spec = (flags & 0x8000u) != 0 && (flags & 0x600u) != 0;
    sbk::g_k1_w64 = (flags & 0x10000u) != 0 && (flags & 0x600u) != 0;
    sbk::g_k1_w64_aligned = (flags & 0x20000u) != 0;
    sbk::g_k1_w64_gt = (flags & 0x80000u) != 0;
    sbk::g_k1_unaligned = (flags & 0x40000u) != 0 && (flags & 0x600u) != 0;      // 64-position step (shared-memory-table chains)   // only meaningful for the multi-chain layouts
    sbk::g_k1_exact = (flags & 0x1000000u) != 0;
    if (flags & 0x1000000u) {
        const unsigned nc = 6, ng = 8;
        std::vector<uint64_t> rings((size_t)grid * (nc + ng) * sbk::K1_RING_GW, 0xCDCDCDCDCDCDCDCDull);
        std::vector<uint16_t> gt((size_t)grid * (ng + 1) * (sbk::K1_TABLE_BYTES / 2) + 8, 0xCDCD);
        uint32_t work = 0;
        a.rings = rings.data();
        a.gtables = (uint16_t*)(((uintptr_t)gt.data() + 15) & ~(uintptr_t)15);
        a.work = &work;
        unsigned chains = (flags >> 20) & 15u;
        if (chains == 0 || chains > nc + ng) chains = nc + ng;
        sbemu::launch(grid, chains * 64, sbk::k1_multi_smem(nc, ng, true), k1_entry, &a);
        return 0;
    }
    if (flags & 0x600u) {
        const unsigned ng = (flags & 0x400u) ? 4 : 0;
        std::vector<uint64_t> rings((size_t)grid * (7 + ng) * sbk::K1_RING_GW, 0xCDCDCDCDCDCDCDCDull);
        std::vector<uint16_t> gt((size_t)grid * (ng + 1) * (sbk::K1_TABLE_BYTES / 2) + 8, 0xCDCD);
        uint32_t work = 0;
        a.rings = rings.data();
        a.gtables = (uint16_t*)(((uintptr_t)gt.data() + 15) & ~(uintptr_t)15);
        a.work = &work;
        unsigned chains = (flags >> 20) & 15u;                  // 0 = all chains of the layout
        if (chains == 0 || chains > 7 + ng) chains = 7 + ng;
        sbemu::launch(grid, chains * 64, 7 * sbk::K1_TABLE_BYTES + (7 + ng) * 64, k1_entry, &a);
        return 0;
    }
    const unsigned np = (flags >> 12) & 7u;
    sbemu::launch(grid, ((np < 1 ? 1 : np > 3 ? 3 : np) + 1) * 32, sbk::K1_SMEM_BYTES, k1_entry, &a);
    return 0;
}

void emu_k1_step_stat(unsigned long* out) { for (int i = 0; i < 3; i++) { out[i] = sbk::g_k1_w32_stat[i]; out[3 + i] = sbk::g_k1_w64_stat[i]; } out[6] = sbk::g_k1_w64_stat[3]; out[7] = sbk::g_k1_w64_stat[4]; }
void emu_k1x_stat(unsigned long* out, int reset) { for (int i = 0; i < 6; i++) { out[i] = sbk::g_k1x_stat[i]; if (reset) sbk::g_k1x_stat[i] = 0; } }
void emu_k1_spec_stat(unsigned long* out) { out[0] = sbk::g_k1_spec_stat[0]; out[1] = sbk::g_k1_spec_stat[1]; }

int emu_decompress_batch(const sb_batch* b, unsigned grid, unsigned block) {
    sb_batch c = *b;
    sbemu::launch(grid, block, (block / 32) * sbk::K2_SMEM_PER_WARP, k2_entry, &c);
    return 0;
}

}
