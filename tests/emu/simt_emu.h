// simt_emu.h -- TEST TOOLING ONLY. A fiber-based emulator of the handful of
// CUDA warp/block primitives the kernel bodies in rust-snappy_b200/csrc use, so
// that the kernels' integer logic can be exercised by g++ on a box with no GPU
// (tests/test_emu_kernels.py). One fiber per CUDA thread; blocks run one after
// another; warp collectives are rendezvous points between the 32 fibers of a
// warp. It checks logic, not memory-model races and not performance.
#pragma once
#define SB_EMU_PRIMITIVES 1
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <vector>

#define SB_DEVICE static inline
#define SB_DEVICE_NOINLINE static
#define SB_FULL 0xFFFFFFFFu

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

namespace sbemu {

struct Fiber;
struct Warp {
    uint64_t buf[2][32];
    int count[2] = {0, 0};
    int nlanes = 32;
};
struct Block {
    std::vector<Fiber*> fibers;
    std::vector<Warp> warps;
    unsigned char* smem = nullptr;
    unsigned block_idx = 0, grid_dim = 1, block_dim = 0;
    int bar_count[2] = {0, 0};
    int named_count[16][2] = {};
    void* sched_sp = nullptr;
};
struct Fiber {
    void* sp = nullptr;
    unsigned char* stack = nullptr;
    Block* blk = nullptr;
    unsigned tid = 0;
    bool done = false;
    unsigned coll_k = 0;      // warp collective generation
    unsigned bar_k = 0;       // block barrier generation
    unsigned named_k[16] = {};
    void (*entry)(void*) = nullptr;
    void* arg = nullptr;
};

extern thread_local Fiber* g_cur;

extern "C" void sbemu_switch(void** save_sp, void* load_sp);
void yield();
void run_block(Block& b, void (*entry)(void*), void* arg, size_t smem_bytes);

// launch <<<grid, block, smem>>> of `entry(arg)`
void launch(unsigned grid, unsigned block, size_t smem_bytes, void (*entry)(void*), void* arg);

// all-gather of a 64-bit value over the calling fiber's warp
inline const uint64_t* warp_gather(uint64_t v) {
    Fiber* f = g_cur;
    Warp& w = f->blk->warps[f->tid >> 5];
    unsigned b = f->coll_k & 1u;
    w.buf[b][f->tid & 31u] = v;
    if (++w.count[b] == w.nlanes) w.count[b ^ 1u] = 0;
    while (w.count[b] < w.nlanes) yield();
    f->coll_k++;
    return w.buf[b];
}

}  // namespace sbemu

namespace sbk {

SB_DEVICE unsigned thread_idx() { return sbemu::g_cur->tid; }
SB_DEVICE unsigned lane_id() { return sbemu::g_cur->tid & 31u; }
SB_DEVICE unsigned warp_id() { return sbemu::g_cur->tid >> 5; }
SB_DEVICE unsigned block_dim() { return sbemu::g_cur->blk->block_dim; }
SB_DEVICE unsigned block_idx() { return sbemu::g_cur->blk->block_idx; }
SB_DEVICE unsigned grid_dim() { return sbemu::g_cur->blk->grid_dim; }

SB_DEVICE uint64_t shfl(uint64_t v, unsigned src) {
    uint64_t r = sbemu::warp_gather(v)[src & 31u];
    return r;
}
SB_DEVICE uint32_t shfl(uint32_t v, unsigned src) { return (uint32_t)shfl((uint64_t)v, src); }
SB_DEVICE int shfl(int v, unsigned src) { return (int)(uint32_t)shfl((uint64_t)(uint32_t)v, src); }
SB_DEVICE uint32_t shfl_up(uint32_t v, unsigned d) {
    unsigned l = lane_id();
    const uint64_t* a = sbemu::warp_gather(v);
    return l >= d ? (uint32_t)a[l - d] : v;
}
SB_DEVICE uint32_t shfl_down(uint32_t v, unsigned d) {
    unsigned l = lane_id();
    const uint64_t* a = sbemu::warp_gather(v);
    return l + d < 32 ? (uint32_t)a[l + d] : v;
}
SB_DEVICE uint32_t shfl_xor(uint32_t v, unsigned m) {
    unsigned l = lane_id();
    return (uint32_t)sbemu::warp_gather(v)[(l ^ m) & 31u];
}
SB_DEVICE uint32_t ballot(bool p) {
    const uint64_t* a = sbemu::warp_gather(p ? 1 : 0);
    uint32_t m = 0;
    for (int i = 0; i < 32; i++) m |= (uint32_t)(a[i] & 1) << i;
    return m;
}
SB_DEVICE bool any(bool p) { return ballot(p) != 0; }
SB_DEVICE bool all(bool p) { return ballot(p) == 0xFFFFFFFFu; }
SB_DEVICE uint32_t match_any(uint32_t v) {
    const uint64_t* a = sbemu::warp_gather(v);
    uint32_t m = 0;
    for (int i = 0; i < 32; i++) if ((uint32_t)a[i] == v) m |= 1u << i;
    return m;
}
SB_DEVICE void syncwarp() { (void)sbemu::warp_gather(0); }
SB_DEVICE void syncthreads() {
    sbemu::Fiber* f = sbemu::g_cur;
    sbemu::Block* b = f->blk;
    unsigned g = f->bar_k & 1u;
    if (++b->bar_count[g] == (int)b->block_dim) b->bar_count[g ^ 1u] = 0;
    while (b->bar_count[g] < (int)b->block_dim) sbemu::yield();
    f->bar_k++;
}
SB_DEVICE void bar_sync(unsigned id, unsigned nthreads) {
    sbemu::Fiber* f = sbemu::g_cur;
    sbemu::Block* b = f->blk;
    unsigned g = f->named_k[id] & 1u;
    if (++b->named_count[id][g] == (int)nthreads) b->named_count[id][g ^ 1u] = 0;
    while (b->named_count[id][g] < (int)nthreads) sbemu::yield();
    f->named_k[id]++;
}

SB_DEVICE void bar_arrive(unsigned id, unsigned nthreads) {
    sbemu::Fiber* f = sbemu::g_cur;
    sbemu::Block* b = f->blk;
    unsigned g = f->named_k[id] & 1u;
    if (++b->named_count[id][g] == (int)nthreads) b->named_count[id][g ^ 1u] = 0;
    f->named_k[id]++;
}

SB_DEVICE int popc(uint32_t v) { return __builtin_popcount(v); }
SB_DEVICE int ffs(uint32_t v) { return __builtin_ffs((int)v); }
SB_DEVICE int clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
SB_DEVICE int ffsll(uint64_t v) { return __builtin_ffsll((long long)v); }
SB_DEVICE uint32_t funnel_r(uint32_t lo, uint32_t hi, unsigned sh) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> (sh & 31u));
}
SB_DEVICE uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s) {
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)(v >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}

SB_DEVICE uint32_t atomic_add(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
SB_DEVICE unsigned long long atomic_add(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
SB_DEVICE uint32_t atomic_min(uint32_t* p, uint32_t v) { uint32_t o = *p; if (v < o) *p = v; return o; }
SB_DEVICE void threadfence() {}
SB_DEVICE void threadfence_block() {}
SB_DEVICE uint32_t reduce_or(uint32_t v) {
    const uint64_t* a = sbemu::warp_gather(v);
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= (uint32_t)a[i];
    return r;
}
SB_DEVICE uint32_t reduce_add(uint32_t v) {
    const uint64_t* a = sbemu::warp_gather(v);
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r += (uint32_t)a[i];
    return r;
}
SB_DEVICE uint32_t reduce_max(uint32_t v) {
    const uint64_t* a = sbemu::warp_gather(v);
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) if ((uint32_t)a[i] > r) r = (uint32_t)a[i];
    return r;
}
SB_DEVICE void spin() { sbemu::yield(); }
SB_DEVICE void spin_long() { sbemu::yield(); }
SB_DEVICE uint32_t ld_volatile(const uint32_t* p) { return *(const volatile uint32_t*)p; }
SB_DEVICE void st_volatile(uint32_t* p, uint32_t v) { *(volatile uint32_t*)p = v; }
SB_DEVICE uint64_t ld_volatile64(const uint64_t* p) { return *(const volatile uint64_t*)p; }

SB_DEVICE uint32_t ldg32(const void* p) { uint32_t v; memcpy(&v, p, 4); return v; }
SB_DEVICE uint4 ldg128(const void* p) { uint4 v; memcpy(&v, p, 16); return v; }
SB_DEVICE uint8_t ldg8(const void* p) { return *(const uint8_t*)p; }
SB_DEVICE void stcs128(void* p, uint4 v) { memcpy(p, &v, 16); }
SB_DEVICE void st8_stream(uint8_t* p, uint8_t v) { *p = v; }
// mbarrier model: the 8 bytes hold the number of completed phases (expected arrival count 1)
SB_DEVICE void mbar_init(uint64_t* bar, unsigned) { *(volatile uint64_t*)bar = 0; }
SB_DEVICE void mbar_arrive(uint64_t* bar) { *(volatile uint64_t*)bar = *(volatile uint64_t*)bar + 1; }
SB_DEVICE bool mbar_try_wait(uint64_t* bar, unsigned parity, unsigned) {
    if (((*(volatile uint64_t*)bar) & 1u) != parity) return true;
    sbemu::yield();
    return ((*(volatile uint64_t*)bar) & 1u) != parity;
}

SB_DEVICE unsigned char* smem() { return sbemu::g_cur->blk->smem; }

}  // namespace sbk
