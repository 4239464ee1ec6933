// k1_exact.cuh -- K1 parser, second generation: exact 32-position windows with the
// table-independent work software-pipelined off the critical path.
//
// Same contract as the first-generation step (k1_eval + k1_finish in k1_compress.cuh): bit-exact with
// reference src/compress.rs:195-317. What changed is where the latency goes:
//
//  * A window starts where the parse stands (lane i = position w+i), so the entry state is always lane 0.
//  * The chain's input bytes travel through a 1 KB shared-memory byte ring (coalesced 128-byte refills issued
//    ~10 windows ahead), so "my 20 bytes", hashes and any candidate within the last ~500 bytes are shared-memory
//    reads, whatever the window alignment.
//  * Candidate evaluation is speculative and pipelined: two aligned 32-position chunks ahead of the parse a stage
//    reads the table slot T' of every position (stale by at most ~4 windows), issues the five candidate-word loads
//    (the L2 round trip that used to sit on the critical path) and, one or two windows later, packs
//    (T', 4-byte equality, match length exact to 15) into a per-position info ring. At parse time a lane re-reads
//    its slot: T == T' (the common case) means the info is current; T != T' means the slot was written after the
//    speculative read, i.e. by a position less than ~130 bytes back, whose bytes are in the byte ring.
//  * In-window dependencies are resolved exactly instead of cutting the window at the first victim: lanes whose
//    hash also occurs at a lower lane of the window ("dynamic" lanes, found with match.any) carry a second
//    precomputed comparison against the nearest such lane, and the walk over the window's events decides per
//    dynamic lane which candidate the serial encoder would have seen (the highest INSERTED lower lane of the same
//    hash, else the table). No restore, no replay; the commit is one predicated store per lane.
//  * Scan runs that leave stride 1 (more than 32 misses in a row) and the block tail go to k1_serial as before.
#pragma once
// included by k1_compress.cuh after the first-generation step (uses K1State, K1Ring, K1Prod, k1_serial, k1_extend)

namespace sbk {

#define K1_HASH(x) (((uint32_t)(x) * 0x1E35A7BDu) >> shift)   // src/compress.rs:522-526

static const uint32_t K1X_RING_WORDS = 256;                  // byte ring: 1 KB, indexed by absolute address
static const uint32_t K1X_INFO_ENTRIES = 256;                // info ring: one u32 per position, indexed by position
static const uint32_t K1X_SCRATCH_BYTES = (K1X_RING_WORDS + K1X_INFO_ENTRIES) * 4;

#if defined(SB_EMU)
static bool g_k1_exact = false;                              // set by the test harness
#define K1_EXACT_ON g_k1_exact
static unsigned long g_k1x_stat[6] = {0, 0, 0, 0, 0, 0};     // windows, bytes, moved windows, dynamic windows, on-demand evals, bails
#elif defined(K1_LEGACY_PARSER)
#define K1_EXACT_ON false
#else
#define K1_EXACT_ON true
#endif

// 16-byte comparison of two byte strings given as five aligned words + bit shift each.
// returns bit0 = first four bytes equal, bits 1..5 = common prefix length in [4,16] (16 = "16 or more"; only
// meaningful when bit0 is set)
SB_DEVICE uint32_t k1x_cmp(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, unsigned ash,
                           uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3, uint32_t b4, unsigned bsh) {
    const uint32_t x0 = funnel_r(a0, a1, ash) ^ funnel_r(b0, b1, bsh);
    const uint32_t x4 = funnel_r(a1, a2, ash) ^ funnel_r(b1, b2, bsh);
    const uint32_t x8 = funnel_r(a2, a3, ash) ^ funnel_r(b2, b3, bsh);
    const uint32_t x12 = funnel_r(a3, a4, ash) ^ funnel_r(b3, b4, bsh);
    const uint32_t l4 = 4 + ((uint32_t)(ffs(x4) - 1) >> 3);
    const uint32_t l8 = 8 + ((uint32_t)(ffs(x8) - 1) >> 3);
    const uint32_t l12 = x12 ? 12 + ((uint32_t)(ffs(x12) - 1) >> 3) : 16;
    const uint32_t L = x4 ? l4 : x8 ? l8 : l12;
    return (x0 == 0 ? 1u : 0u) | (L << 1);
}

struct K1xPend {            // one chunk between its two pipeline stages
    uint32_t b0, b1, b2, b3, b4;   // candidate words (in flight)
    uint32_t T;                    // the slot value they were fetched for
    uint32_t m;                    // chunk index (positions 32m .. 32m+31)
    bool live;
};

// Parser warp of one chain. `in` = the unit's input in global memory (any alignment), n >= 17.
// `scratch` = K1X_SCRATCH_BYTES of shared memory private to this chain.
template <bool GT>
SB_DEVICE void k1_parse_x(const uint8_t* in, uint32_t n, uint16_t* table, const K1Ring& ring, uint32_t* ctrl,
                          uint32_t* scratch) {
    const unsigned lane = lane_id();
    unsigned shift = 24;
    uint32_t tsize = 256;
    while (tsize < 16384 && tsize < n) { shift--; tsize *= 2; }   // src/compress.rs:491-497
    const uint32_t s_limit = n - 15;
    K1Prod prod;
    prod.head = ld_volatile(&ctrl[6]); prod.published = ld_volatile(&ctrl[7]); prod.tail_seen = 0;
    K1State st;
    st.s = 1; st.skip = 32; st.rematch = false;
    uint32_t* const rw = scratch;
    uint32_t* const iw = scratch + K1X_RING_WORDS;
    const uintptr_t base = (uintptr_t)in;
    const uintptr_t first_word = base & ~(uintptr_t)3, endA = base + n;
    // ---- byte ring: every address in [hi - 1024, hi) that lies inside the block is present
    uintptr_t hi = base & ~(uintptr_t)127;
    uint32_t pf = 0;
    bool pf_live = false;
    auto ring_load = [&](uintptr_t a) -> uint32_t {
        const uintptr_t wa = a + 4 * lane;
        return (wa >= first_word && wa < endA) ? *(const uint32_t*)wa : 0u;
    };
    // ---- chunk pipeline: F chunks complete (info valid), I issued; pa = older pending set, pb = newer
    uint32_t F = 0, I = 0;
    K1xPend pa, pb;
    pa.b0 = pa.b1 = pa.b2 = pa.b3 = pa.b4 = pa.T = pa.m = 0; pa.live = false;
    pb = pa;
    auto issue = [&](uint32_t m, K1xPend& P) {
        const uintptr_t A = base + 32u * m + lane;
        const uint32_t wi = (uint32_t)(A >> 2);
        const uint32_t cur = funnel_r(rw[wi & 255u], rw[(wi + 1) & 255u], (unsigned)(A & 3u) * 8);
        const uint32_t T = table[K1_HASH(cur)];
        const uint32_t* bw = (const uint32_t*)((base + T) & ~(uintptr_t)3);
        P.b0 = bw[0]; P.b1 = bw[1]; P.b2 = bw[2]; P.b3 = bw[3]; P.b4 = bw[4];
        P.T = T; P.m = m; P.live = true;
    };
    auto complete = [&](const K1xPend& P) {
        const uintptr_t A = base + 32u * P.m + lane;
        const uint32_t wi = (uint32_t)(A >> 2);
        const uint32_t r = k1x_cmp(rw[wi & 255u], rw[(wi + 1) & 255u], rw[(wi + 2) & 255u], rw[(wi + 3) & 255u],
                                   rw[(wi + 4) & 255u], (unsigned)(A & 3u) * 8,
                                   P.b0, P.b1, P.b2, P.b3, P.b4, (unsigned)((base + P.T) & 3u) * 8);
        iw[(32u * P.m + lane) & 255u] = P.T | (r << 16);
    };

    K1_PROF_DECL
    for (;;) {
        const uint32_t w = st.s;
        K1_TICK(8);                                                  // [8] exit state / publish / loop
        // fast path: every lane of the window is an ordinary probe position (w+31 < s_limit with room for the
        // 20-byte reads of the chunks that cover it) and the scan stride is 1
        const bool fast = w + 90 < n && (st.rematch || st.skip < 64);
        if (!fast) {
            bool finished;
            if (st.rematch ? st.s >= s_limit : st.s + (st.skip >> 5) > s_limit) finished = true;
            else finished = k1_serial(in, n, table, shift, s_limit, st, w + 32, ring, prod);
            K1_TICK(9);                                              // [9] serial path
            if (finished) break;
            continue;
        }
        // ---- byte ring upkeep
        const uintptr_t Aw = base + w;
        if (Aw > hi + 256) { hi = (Aw - 256) & ~(uintptr_t)127; pf_live = false; }   // far jump: restart the ring behind the window
        if (pf_live) { rw[((uint32_t)(hi >> 2) + lane) & 255u] = pf; hi += 128; pf_live = false; }
        while (hi < Aw + 160) { rw[((uint32_t)(hi >> 2) + lane) & 255u] = ring_load(hi); hi += 128; }
        if (hi + 128 <= Aw + 512) { pf = ring_load(hi); pf_live = true; }
        syncwarp();
        K1_TICK(0);                                                  // [0] byte ring upkeep
        // ---- chunk pipeline upkeep
        const uint32_t c0 = w >> 5, need = ((w + 31) >> 5) + 1;
        if (c0 > I) { I = F = c0; pa.live = false; pb.live = false; }
        while (F < c0) { pa = pb; pb.live = false; F++; }          // pending chunks the parse jumped over
        while (F < need) {
            if (I == F) { issue(I, pa); I++; }
            complete(pa);
            pa = pb; pb.live = false; F++;
        }
        while (I - F < 2 && I < need + 2 && 32u * I + 56 < n) {
            if (!pa.live) issue(I, pa); else issue(I, pb);
            I++;
        }
        syncwarp();
        K1_TICK(1);                                                  // [1] chunk pipeline (complete + issue)
        k1_wait_space(ring, prod, 12);                              // a window holds at most 9 copies
        K1_TICK(2);                                                  // [2] event ring space
        // ---- this window: own bytes, hash, current slot, speculative info
        const uintptr_t A = Aw + lane;
        const uint32_t wi = (uint32_t)(A >> 2);
        const unsigned ash = (unsigned)(A & 3u) * 8;
        const uint32_t a0 = rw[wi & 255u], a1 = rw[(wi + 1) & 255u], a2 = rw[(wi + 2) & 255u], a3 = rw[(wi + 3) & 255u],
                       a4 = rw[(wi + 4) & 255u];
        const uint32_t h = K1_HASH(funnel_r(a0, a1, ash));
        const uint32_t T = table[h];
        const uint32_t info = iw[(w + lane) & 255u];
        uint32_t pk0 = info >> 16;                                   // bit0 eq, bits 1..5 length
        const bool moved = (info & 0xFFFFu) != T;
        K1_TICK(3);                                                  // [3] own bytes, hash, slot, info
        if (any(moved)) {
            // the slot was written after the speculative read: T is less than ~130 bytes back, in the byte ring
            const uintptr_t B = base + T;
            const uint32_t bi = (uint32_t)(B >> 2);
            const uint32_t r = k1x_cmp(a0, a1, a2, a3, a4, ash, rw[bi & 255u], rw[(bi + 1) & 255u], rw[(bi + 2) & 255u],
                                       rw[(bi + 3) & 255u], rw[(bi + 4) & 255u], (unsigned)(B & 3u) * 8);
            if (moved) pk0 = r;
        }
        K1_TICK(4);                                                  // [4] moved slots re-evaluated
        // dynamic lanes: a lower lane of this window has the same hash
        const uint32_t same = match_any(h);
        const uint32_t below = same & ((1u << lane) - 1u);
        const uint32_t D = ballot(below != 0);
        uint32_t pk1 = 0, pin = 0;
        if (D) {
            pin = below ? 31u - (uint32_t)clz(below) : 0u;
            const uintptr_t B = Aw + pin;
            const uint32_t bi = (uint32_t)(B >> 2);
            pk1 = k1x_cmp(a0, a1, a2, a3, a4, ash, rw[bi & 255u], rw[(bi + 1) & 255u], rw[(bi + 2) & 255u],
                          rw[(bi + 3) & 255u], rw[(bi + 4) & 255u], (unsigned)(B & 3u) * 8);
        }
        const uint32_t pk = pk0 | (pk1 << 6) | (pin << 12);          // [5:0] vs table, [11:6] vs pin, [16:12] pin
        const uint32_t EV = ballot((pk0 & 1u) != 0 && below == 0) | D;   // static hits + every dynamic lane
#if defined(SB_EMU)
        { const bool am = any(moved); if (lane == 0) { g_k1x_stat[0]++; g_k1x_stat[2] += am ? 1 : 0; g_k1x_stat[3] += D ? 1 : 0; } }
#endif
#ifdef SB_EMU_TRACE
        if (lane == 0) fprintf(stderr, "x win w=%u rm=%d skip=%u EV=%08x D=%08x F=%u I=%u\n", w, (int)st.rematch, st.skip, EV, D, F, I);
#endif
#ifdef SB_EMU_TRACE
        if (w == 1 && lane < 4) fprintf(stderr, "x   lane %u h=%u T=%u info=%08x moved=%d pk0=%x pk1=%x pin=%u below=%x pk=%x\n", lane, h, T, info, (int)moved, pk0, pk1, pin, below, pk);
#endif
        K1_TICK(5);                                                  // [5] match.any + dynamic lanes
        // ---- walk the events of the window in stream order (all lanes compute the same values)
        uint32_t INS = 0;                 // lanes inserted so far (probes and copy-end pre-inserts)
        uint32_t pos = 0;                 // next probe (lane index); rm: it is the probe right after a copy
        bool rm = st.rematch;
        uint32_t skip = st.skip;
        bool bail = false, finished = false;
        uint32_t far_pre = 0;             // absolute position of a copy-end pre-insert beyond the window (0 = none)
        while (pos < 32) {
            const uint32_t m = EV >> pos;
            const uint32_t f = m ? pos + (uint32_t)(ffs(m) - 1) : 31u;
            const uint32_t cnt = f - pos + 1;
            if (!rm && skip + cnt > 64) { bail = true; break; }      // the run leaves stride 1 inside the window
            INS |= (cnt >= 32 ? 0xFFFFFFFFu : ((1u << cnt) - 1u)) << pos;
            const uint32_t skip_after = rm ? 32 + cnt - 1 : skip + cnt;
            bool hit = false;
            uint32_t L = 0, q = 32;       // q: in-window candidate lane (32 = the table's candidate)
            if (m) {
                const uint32_t pkf = shfl(pk, f);
                hit = (pkf & 1u) != 0; L = (pkf >> 1) & 31u;
                if ((D >> f) & 1u) {
                    const uint32_t qm = shfl(below, f) & INS;         // inserted lower lanes of the same hash
                    if (qm) {
                        q = 31u - (uint32_t)clz(qm);
                        if (q == ((pkf >> 12) & 31u)) { hit = ((pkf >> 6) & 1u) != 0; L = (pkf >> 7) & 31u; }
                        else {
                            // the nearest same-hash lane was not inserted but an older one was: evaluate on demand
                            const uintptr_t B = Aw + q;
                            const uint32_t bi = (uint32_t)(B >> 2);
                            const uint32_t r = shfl(k1x_cmp(a0, a1, a2, a3, a4, ash, rw[bi & 255u], rw[(bi + 1) & 255u],
                                                            rw[(bi + 2) & 255u], rw[(bi + 3) & 255u], rw[(bi + 4) & 255u],
                                                            (unsigned)(B & 3u) * 8), f);
                            hit = (r & 1u) != 0; L = (r >> 1) & 31u;
#if defined(SB_EMU)
                            if (lane == 0) g_k1x_stat[4]++;
#endif
                        }
                    }
                }
            }
#ifdef SB_EMU_TRACE
            if (lane == 0) fprintf(stderr, "x   ev pos=%u f=%u cnt=%u hit=%d L=%u q=%u INS=%08x\n", pos, f, cnt, (int)hit, L, q, INS);
#endif
            if (!hit) { skip = skip_after; rm = false; pos = f + 1; continue; }
            // ---- copy at lane f (:258-276)
            const uint32_t p = w + f;
            const uint32_t cand = q < 32 ? w + q : shfl(T, f);
            if (L >= 16) L = k1_extend(in, n, p + 16, cand + 16) - p;
#ifdef SB_EMU_TRACE
            if (lane == 0) fprintf(stderr, "x copy w=%u f=%u p=%u len=%u cand=%u q=%u\n", w, f, p, L, cand, q);
#endif
            if (lane == 0) ring.ev[prod.head & (ring.size - 1)] = k1_event(p, L, p - cand);
            prod.head++;
            const uint32_t e = f + L;
            if (w + e >= s_limit) { finished = true; break; }
            if (e <= 32) INS |= 1u << (e - 1);                        // pre-insert of e-1 (:293-295)
            else far_pre = w + e - 1;
            pos = e; rm = true;
        }
        K1_TICK(6);                                                  // [6] walk
        // ---- commit: the highest inserted lane of every hash owns the slot
        if (!finished) {
            syncwarp();
            if (((INS >> lane) & 1u) && (same & INS & ~((2u << lane) - 1u)) == 0) table[h] = (uint16_t)(w + lane);
            syncwarp();
            if (far_pre) {
                if (far_pre + 8 < w + 150) {
                    const uintptr_t B = base + far_pre;
                    const uint32_t bi = (uint32_t)(B >> 2);
                    const uint32_t hh = K1_HASH(funnel_r(rw[bi & 255u], rw[(bi + 1) & 255u], (unsigned)(B & 3u) * 8));
                    if (lane == 0) table[hh] = (uint16_t)far_pre;
                    syncwarp();
                } else {
                    k1_preinsert(in, table, shift, s_limit, far_pre + 1);
                }
            }
        }
        K1_TICK(7);                                                  // [7] commit
        if (prod.head - prod.published >= K1_PUBLISH) k1_publish(ring, prod);
#if defined(K1_PROFILE) && defined(__CUDACC__)
        k1_acc[11]++;                                                // windows
#endif
#if defined(SB_EMU)
        if (lane == 0) { g_k1x_stat[1] += (finished ? 0 : w + pos - st.s); g_k1x_stat[5] += bail ? 1 : 0; }
#endif
        if (finished) break;
        st.s = w + pos; st.rematch = rm; st.skip = skip;
        if (bail) {
            // stride > 1 inside the window: the reference's control flow takes over from the accepted prefix
            if (k1_serial(in, n, table, shift, s_limit, st, w + 32, ring, prod)) break;
        }
    }
    k1_push(ring, prod, k1_event(n, 0, 0));                          // end marker -> trailing literal (:417-426)
    k1_publish(ring, prod);
    K1_TICK(10);
    K1_PROF_FLUSH;
    syncwarp();
    if (lane == 0) { ctrl[6] = prod.head; ctrl[7] = prod.published; }
}

#undef K1_HASH

}  // namespace sbk
