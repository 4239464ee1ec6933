// emu_kernels.cpp -- TEST TOOLING ONLY. Compiles the CUDA kernel bodies of
// rust-snappy_b200/csrc with g++ against the fiber warp emulator and exposes
// them to pytest through a C interface (tests/test_emu_kernels.py).
#define SB_EMU 1
#include "../../rust-snappy_b200/csrc/k1_compress.cuh"
#include "../../rust-snappy_b200/csrc/k2_decompress.cuh"

struct K1Args { sb_batch b; uint32_t flags; };
static void k1_entry(void* a) { K1Args* x = (K1Args*)a; sbk::k1_compress_body(x->b, x->flags); }
static void k2_entry(void* a) { sbk::k2_decompress_body(*(sb_batch*)a); }

extern "C" {

int emu_compress_batch(const sb_batch* b, uint32_t flags, unsigned grid) {
    K1Args a{*b, flags};
    sbemu::launch(grid, sbk::K1_THREADS, sbk::K1_SMEM_BYTES, k1_entry, &a);
    return 0;
}

int emu_decompress_batch(const sb_batch* b, unsigned grid, unsigned block) {
    sb_batch c = *b;
    sbemu::launch(grid, block, 0, k2_entry, &c);
    return 0;
}

}
