// simt_emu.cpp -- TEST TOOLING ONLY (see simt_emu.h).
#include "simt_emu.h"

namespace sbemu {

thread_local Fiber* g_cur = nullptr;
static thread_local Block* g_blk = nullptr;

// Minimal x86-64 SysV context switch: callee-saved registers + stack pointer.
__asm__(
    ".text\n"
    ".globl sbemu_switch\n"
    ".type sbemu_switch,@function\n"
    "sbemu_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size sbemu_switch,.-sbemu_switch\n");

void yield() {
    Fiber* f = g_cur;
    sbemu_switch(&f->sp, f->blk->sched_sp);
}

static void trampoline() {
    Fiber* f = g_cur;
    f->entry(f->arg);
    f->done = true;
    for (;;) yield();
}

static const size_t STACK_BYTES = 256 * 1024;

void run_block(Block& b, void (*entry)(void*), void* arg, size_t smem_bytes) {
    b.smem = (unsigned char*)aligned_alloc(128, (smem_bytes + 127 + 128) / 128 * 128);
    memset(b.smem, 0xCD, smem_bytes);  // poison: kernels must initialise what they read
    unsigned nthreads = b.block_dim;
    b.warps.assign((nthreads + 31) / 32, Warp());
    for (unsigned w = 0; w < b.warps.size(); w++) {
        unsigned lanes = nthreads - w * 32;
        b.warps[w].nlanes = lanes > 32 ? 32 : (int)lanes;
    }
    std::vector<Fiber> fibers(nthreads);
    for (unsigned t = 0; t < nthreads; t++) {
        Fiber& f = fibers[t];
        f.blk = &b; f.tid = t; f.entry = entry; f.arg = arg;
        f.stack = (unsigned char*)aligned_alloc(64, STACK_BYTES);
        // initial frame: 6 callee-saved slots + return address -> trampoline
        uintptr_t top = ((uintptr_t)(f.stack + STACK_BYTES)) & ~(uintptr_t)15;
        void** sp = (void**)top;
        *--sp = nullptr;                 // alignment slot (so that rsp%16==8 at entry)
        *--sp = (void*)&trampoline;      // return address for `ret`
        for (int i = 0; i < 6; i++) *--sp = nullptr;
        f.sp = (void*)sp;
    }
    g_blk = &b;
    unsigned live = nthreads;
    // SBEMU_ORDER=reverse|shuffle perturbs the lane schedule to flush out code that
    // silently relies on lanes running in index order
    const char* ord = getenv("SBEMU_ORDER");
    const int mode = !ord ? 0 : (ord[0] == 'r' ? 1 : 2);
    unsigned rng = 12345u + b.block_idx;
    while (live) {
        live = 0;
        for (unsigned k = 0; k < nthreads; k++) {
            unsigned t = k;
            if (mode == 1) t = nthreads - 1 - k;
            else if (mode == 2) { rng = rng * 1664525u + 1013904223u; t = (k + (rng >> 16)) % nthreads; }
            Fiber& f = fibers[t];
            if (f.done) continue;
            g_cur = &f;
            sbemu_switch(&b.sched_sp, f.sp);
        }
        for (unsigned t = 0; t < nthreads; t++) live += fibers[t].done ? 0 : 1;
    }
    g_cur = nullptr;
    for (auto& f : fibers) free(f.stack);
    free(b.smem);
    b.smem = nullptr;
}

void launch(unsigned grid, unsigned block, size_t smem_bytes, void (*entry)(void*), void* arg) {
    for (unsigned bi = 0; bi < grid; bi++) {
        Block b;
        b.block_idx = bi; b.grid_dim = grid; b.block_dim = block;
        run_block(b, entry, arg, smem_bytes);
    }
}

}  // namespace sbemu
