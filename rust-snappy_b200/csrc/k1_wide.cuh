// k1_wide.cuh -- EXPERIMENTAL 64-position parser step for K1 (two positions per lane).
//
// Not part of the default product path: compiled in with -DK1_W64 (tools/build_variant.sh w64 -DK1_W64), exercised by
// the emulator tests, not measured on hardware yet. Motivation (tools/sim_window_width.py): the
// per-step costs of the parser (loop top, the L2 round trip of the candidate words, commit, event ring, exit state)
// are paid once per step, and text resolves 52 bytes per 64-position step against 30 per 32-position step.
//
// Same contract as k1_eval + k1_finish (k1_compress.cuh): bit-exact with src/compress.rs:195-317.
//   lane l owns positions w+l ("half 0") and w+32+l ("half 1"); masks over the window are 64-bit.
//   * probe both positions against the table as of the step start;
//   * walk the taken copies from the entry state, one hop per copy (all lanes compute the same values), extending
//     >= 16-byte matches cooperatively on the way and checking that no scan run inside the window leaves stride 1;
//   * inserted positions = entry..63 minus copy interiors; commit half 0, re-read half 1's slots (a slot that
//     moved means that lane should have seen a candidate from inside this step: a "victim", the step is cut
//     there), commit half 1; same-slot clashes inside a half are resolved as in k1_finish;
//   * publish the copy events, leave the exit state.
// Only used by chains whose table lives in shared memory (the slot re-read is cheap there).
#pragma once
// included by k1_compress.cuh (after the 32-position step it builds on)

namespace sbk {

#define K1_HASH(x) (((uint32_t)(x) * 0x1E35A7BDu) >> shift)   // src/compress.rs:522-526

#if defined(SB_EMU)
static bool g_k1_w64 = false;                          // set by the test harness
static bool g_k1_w64_aligned = false;                  // ... windows on 64-byte boundaries instead of starting at the parse position
#define K1_W64_ALIGNED_ON g_k1_w64_aligned
static unsigned long g_k1_w64_stat[5] = {0, 0, 0, 0, 0};   // fast steps, bytes resolved, serial replays, clash resolutions, cuts
#define K1_W64_ON g_k1_w64
#elif defined(K1_W64)
#define K1_W64_ON true
#else
#define K1_W64_ON false
#endif
// -DK1_W64_GT: the chains whose table lives in L2 take 64-position steps too; their commit finds slot clashes and
// victims by comparing hashes across lanes (match.any) BEFORE storing, so it never reads the table back
#if defined(SB_EMU)
static bool g_k1_w64_gt = false;
#define K1_W64_GT_ON g_k1_w64_gt
#elif defined(K1_W64_GT)
#define K1_W64_GT_ON true
#else
#define K1_W64_GT_ON false
#endif
#if !defined(SB_EMU)
#if defined(K1_W64_ALIGNED)
#define K1_W64_ALIGNED_ON true
#else
#define K1_W64_ALIGNED_ON false
#endif
#endif

struct K1Seq64 { uint32_t a[2][5]; uint32_t w; };
SB_DEVICE K1Seq64 k1_fetch_seq64(const uint8_t* win, uint32_t w) {
    const uintptr_t aa = (uintptr_t)(win + w + lane_id());
    const uint32_t* aw = (const uint32_t*)(aa & ~(uintptr_t)3);
    K1Seq64 q;
#pragma unroll
    for (int k = 0; k < 5; k++) { q.a[0][k] = aw[k]; q.a[1][k] = aw[8 + k]; }
    q.w = w;
    return q;
}

struct K1Pre64 {
    uint32_t h[2], c[2], L[2];   // per half: hash, candidate, match length (exact to 15, 16 = "16 or more")
    bool eq[2];
    uint64_t E, longs;           // hit mask / hits whose length is only known to be >= 16
};

SB_DEVICE uint64_t k1_below64(uint32_t k) { return k >= 64 ? ~0ull : ((1ull << k) - 1ull); }
SB_DEVICE uint64_t k1_ballot64(bool p0, bool p1) { return (uint64_t)ballot(p0) | ((uint64_t)ballot(p1) << 32); }
// value of entry q (0..63) of a per-half pair; q is warp-uniform
SB_DEVICE uint32_t k1_pick64(uint32_t v0, uint32_t v1, uint32_t q) { return q < 32 ? shfl(v0, q) : shfl(v1, q - 32); }

SB_DEVICE K1Pre64 k1_eval64(const uint8_t* win, const uint16_t* table, unsigned shift, uint32_t w, const K1Seq64& seq) {
    K1Pre64 r;
    const unsigned ash = (unsigned)((uintptr_t)(win + w + lane_id()) & 3u) * 8;   // both halves share the alignment
    K1Seq64 q = seq;
    if (q.w != w) q = k1_fetch_seq64(win, w);
#pragma unroll
    for (int x = 0; x < 2; x++) {
        const uint32_t a0 = q.a[x][0], a1 = q.a[x][1], a2 = q.a[x][2], a3 = q.a[x][3], a4 = q.a[x][4];
        const uint32_t cur = funnel_r(a0, a1, ash);
        r.h[x] = K1_HASH(cur);
        r.c[x] = table[r.h[x]];
        const uintptr_t ba = (uintptr_t)(win + r.c[x]);
        const uint32_t* bw = (const uint32_t*)(ba & ~(uintptr_t)3);
        const unsigned bsh = (unsigned)(ba & 3u) * 8;
        const uint32_t b0 = bw[0], b1 = bw[1], b2 = bw[2], b3 = bw[3], b4 = bw[4];
        r.eq[x] = cur == funnel_r(b0, b1, bsh);
        const uint32_t x4 = funnel_r(a1, a2, ash) ^ funnel_r(b1, b2, bsh);
        const uint32_t x8 = funnel_r(a2, a3, ash) ^ funnel_r(b2, b3, bsh);
        const uint32_t x12 = funnel_r(a3, a4, ash) ^ funnel_r(b3, b4, bsh);
        const uint32_t l4 = 4 + ((uint32_t)(ffs(x4) - 1) >> 3);
        const uint32_t l8 = 8 + ((uint32_t)(ffs(x8) - 1) >> 3);
        const uint32_t l12 = x12 ? 12 + ((uint32_t)(ffs(x12) - 1) >> 3) : 16;
        r.L[x] = x4 ? l4 : x8 ? l8 : l12;
    }
    r.E = k1_ballot64(r.eq[0], r.eq[1]);
    r.longs = k1_ballot64(r.eq[0] && r.L[0] == 16, r.eq[1] && r.L[1] == 16);
    return r;
}

// Commit the inserts of one half. `ins`: lanes of this half that insert; `pre`: lanes that are write-only copy-end
// inserts (never victims); `old`: the slot value to put back when the commit has to be redone; `limit`: lanes at or
// above it are not part of the step (a victim found earlier). Returns the cut: the first victim lane of this half
// (`limit` if none); on return the table holds exactly the inserts of lanes below the cut, last writer of a slot wins.
// `peek_h` / `peek`: a slot every lane wants to read as of AFTER this commit (issued together with the verify read).
SB_DEVICE uint32_t k1_commit_half(uint16_t* table, uint32_t h, uint32_t p, uint32_t old, uint32_t ins, uint32_t pre,
                                  uint32_t limit, uint32_t peek_h, uint32_t& peek) {
    const unsigned lane = lane_id();
    const uint32_t keep0 = ins & (limit >= 32 ? 0xFFFFFFFFu : ((1u << limit) - 1u));
    const bool my = (keep0 >> lane) & 1u;
    syncwarp();                                                  // every probe read precedes the commit
    if (my) table[h] = (uint16_t)p;
    syncwarp();
    const uint32_t seen = table[h];
    peek = table[peek_h];
    const bool clash = my && seen != (uint16_t)p;
    uint32_t cut = limit;
    if (any(clash)) {
#if defined(SB_EMU)
        if (lane == 0) g_k1_w64_stat[3]++;
#endif
        const uint32_t same = match_any(my ? h : 0xFFFF0000u | lane);
        const bool victim = my && !((pre >> lane) & 1u) && (same & keep0 & ((1u << lane) - 1u)) != 0;
        const uint32_t vm = ballot(victim);
        if (vm) cut = ffs(vm) - 1;
        syncwarp();
        if (my) table[h] = (uint16_t)old;
        syncwarp();
        const uint32_t keep = keep0 & (cut >= 32 ? 0xFFFFFFFFu : ((1u << cut) - 1u));
        if (((keep >> lane) & 1u) && (same & keep & ~((2u << lane) - 1u)) == 0) table[h] = (uint16_t)p;
        syncwarp();
        peek = table[peek_h];
    }
    return cut;
}

// Commit of a 64-position step without reading the table back (for tables in L2): same-slot groups come from
// match.any over the hashes, so victims and the cut are known before anything is stored.
//   C / PRE: inserted positions / write-only copy-end inserts of the step. Returns the cut (64 = whole step accepted);
//   on return the table holds exactly the inserts below the cut, last writer of a slot wins.
SB_DEVICE uint32_t k1_commit64_match(uint16_t* table, uint32_t h0, uint32_t h1, uint32_t p0, uint32_t p1, uint64_t C, uint64_t PRE) {
    const unsigned lane = lane_id();
    const uint32_t C0 = (uint32_t)C, C1 = (uint32_t)(C >> 32), PRE0 = (uint32_t)PRE, PRE1 = (uint32_t)(PRE >> 32);
    const bool ins0 = (C0 >> lane) & 1u, ins1 = (C1 >> lane) & 1u;
    const uint32_t below = (1u << lane) - 1u, above = ~((2u << lane) - 1u);
    const uint32_t key0 = ins0 ? h0 : (0xFFFF0000u | lane);      // hashes are < 16384: these never match anything
    // ---- half 0: a probed lane with a lower inserted lane of the same hash is a victim
    const uint32_t same0 = match_any(key0);
    const uint32_t v0 = ballot(ins0 && !((PRE0 >> lane) & 1u) && (same0 & C0 & below) != 0);
    syncwarp();                                                  // every probe read precedes the commit
    if (v0) {
        const uint32_t cut0 = ffs(v0) - 1;
        const uint32_t keep0 = C0 & ((1u << cut0) - 1u);
        if (((keep0 >> lane) & 1u) && (same0 & keep0 & above) == 0) table[h0] = (uint16_t)p0;
        syncwarp();
        return cut0;
    }
    // ---- half 1 against half 0's inserts: every (i, j) pair has to sit on two different lanes of one match.any,
    // which takes four calls (lanes 0-15 / 16-31 of each side, the second pair with half 1's hashes rotated by 16)
    const bool lo16 = lane < 16;
    const uint32_t q1s = shfl(h1, lane ^ 16u);
    const uint32_t mA = match_any(lo16 ? key0 : h1);             // j in 0..15 , i in 16..31 (on its own lane)
    const uint32_t mB = match_any(lo16 ? h1 : key0);             // j in 16..31, i in 0..15  (on its own lane)
    const uint32_t mC = match_any(lo16 ? key0 : q1s);            // j in 0..15 , i in 0..15  (on lane i+16)
    const uint32_t mD = match_any(lo16 ? q1s : key0);            // j in 16..31, i in 16..31 (on lane i-16)
    const uint32_t moved = ballot(!lo16 && (mA & 0x0000FFFFu) != 0) | ballot(lo16 && (mB & 0xFFFF0000u) != 0) |
                           (ballot(!lo16 && (mC & 0x0000FFFFu) != 0) >> 16) | (ballot(lo16 && (mD & 0xFFFF0000u) != 0) << 16);
    // ---- half 1 itself
    const uint32_t key1 = ins1 ? h1 : (0xFFFF0000u | lane);
    const uint32_t same1 = match_any(key1);
    const uint32_t probed1 = C1 & ~PRE1;
    const uint32_t v1 = ballot(ins1 && !((PRE1 >> lane) & 1u) && (same1 & C1 & below) != 0) | (moved & probed1);
    const uint32_t cut1 = v1 ? (uint32_t)(ffs(v1) - 1) : 32;
    if (ins0 && (same0 & C0 & above) == 0) table[h0] = (uint16_t)p0;
    syncwarp();                                                  // half 1's stores land after half 0's
    const uint32_t keep1 = C1 & (cut1 >= 32 ? 0xFFFFFFFFu : ((1u << cut1) - 1u));
    if (((keep1 >> lane) & 1u) && (same1 & keep1 & above) == 0) table[h1] = (uint16_t)p1;
    syncwarp();
    return 32 + cut1;
}

// One 64-position step from a current probe. Returns false (state and table untouched) when the window has to be
// replayed serially (a scan run inside it leaves stride 1).
template <bool GT>
SB_DEVICE bool k1_finish64(const uint8_t* win, uint32_t n, uint16_t* table, unsigned shift, uint32_t s_limit,
                           K1State& st, const K1Ring& ring, K1Prod& head, const K1Pre64& pre, const K1Seq64& nxt, uint32_t w) {
    const unsigned lane = lane_id();
    const uint32_t i0 = st.s - w;                                // 0 unless windows are kept on 64-byte boundaries
    const uint64_t E = pre.E;
    uint32_t L0 = pre.L[0], L1 = pre.L[1];
    // ---- entry: rematch probe at i0 or scan from i0 -> first hit at/after i0
    const uint64_t fm = E >> i0;
    const uint32_t f = fm ? i0 + (uint32_t)(ffsll(fm) - 1) : 64;
    if (st.rematch) {
        // a miss at i0 starts a scan run at i0+1 with skip 32: it must reach its hit (or the window end) at stride 1
        if (f != i0 && (f < 64 ? f - i0 : 63 - i0) > 32) return false;
    } else {
        const uint32_t probes = f < 64 ? f - i0 + 1 : 64 - i0;
        if (st.skip + probes > 64) return false;
    }
    // ---- per entry, in parallel: where the parse goes after a copy taken here (next hit at/after its end, 64 = leaves
    // the window or no more hits), whether the scan run behind it leaves stride 1, whether its length is still open
    auto hop = [&](uint32_t pos, uint32_t len, bool lng) -> uint32_t {
        const uint32_t e = pos + len;
        uint32_t nx = 64, viol = 0;
        if (e < 64) {
            const uint64_t m = E >> e;
            if (m) nx = e + (uint32_t)(ffsll(m) - 1);
            viol = (nx != e && (nx < 64 ? nx - e : 63 - e) > 32) ? 1u : 0u;
        }
        return nx | (viol << 7) | ((lng ? 1u : 0u) << 8);
    };
    uint32_t pk0 = hop(lane, L0, (pre.longs >> lane) & 1ull), pk1 = hop(32 + lane, L1, (pre.longs >> (32 + lane)) & 1ull);
    // ---- walk the taken copies: one shuffle per hop (every lane follows the same chain)
    uint64_t CS = 0;
    uint32_t last = 64;
    for (uint32_t cur = f; cur < 64;) {
        uint32_t pk = k1_pick64(pk0, pk1, cur);
        if (pk & 0x100u) {                                       // >= 16 bytes: extend cooperatively to the exact end
            const uint32_t cj = k1_pick64(pre.c[0], pre.c[1], cur);
            const uint32_t pj = w + cur;
            const uint32_t Lc = k1_extend(win, n, pj + 16, cj + 16) - pj;
            pk = hop(cur, Lc, false);
            if (lane == (cur & 31u)) { if (cur < 32) { L0 = Lc; pk0 = pk; } else { L1 = Lc; pk1 = pk; } }
        }
        if (pk & 0x80u) return false;                            // the scan run after this copy leaves stride 1
        CS |= 1ull << cur;
        last = cur;
        cur = pk & 0x7Fu;
    }
    const uint32_t e_last = last < 64 ? last + k1_pick64(L0, L1, last) : 0;
    // ---- inserted positions = entry..63 minus copy interiors [q+1, e-2]
    const bool t0 = (CS >> lane) & 1ull, t1 = (CS >> (32 + lane)) & 1ull;
    uint64_t interior = 0;
    if (t0) interior |= k1_below64(lane + L0 - 1) & ~k1_below64(lane + 1);
    if (t1) interior |= k1_below64(32 + lane + L1 - 1) & ~k1_below64(32 + lane + 1);
    const uint64_t I = (uint64_t)reduce_or((uint32_t)interior) | ((uint64_t)reduce_or((uint32_t)(interior >> 32)) << 32);
    const uint64_t C = (~0ull << i0) & ~I;
    // copy-end inserts (e-1) are write-only; e-1 is the position right after a copy's interior (which is never empty)
    const uint64_t PRE = (I << 1) & ~I;
    const uint32_t p0 = w + lane, p1 = w + 32 + lane;
#ifdef SB_EMU_TRACE
    if (lane == 0) fprintf(stderr, "win64 w=%u i0=%u rm=%d skip=%u E=%016llx f=%u CS=%016llx C=%016llx\n", w, i0, (int)st.rematch, st.skip,
                           (unsigned long long)E, f, (unsigned long long)CS, (unsigned long long)C);
#endif
    // ---- commit half 0, find half 1's victims of half 0's inserts, commit half 1
    uint32_t cut;
    if (GT) {
        cut = k1_commit64_match(table, pre.h[0], pre.h[1], p0, p1, C, PRE);
    } else {
        uint32_t c1, unused;
        cut = k1_commit_half(table, pre.h[0], p0, pre.c[0], (uint32_t)C, (uint32_t)PRE, 32, pre.h[1], c1);
        if (cut >= 32) {
            // c1 = half 1's slots after half 0's commit: a probed lane whose slot moved should have seen a candidate
            // from half 0 -> victim
            const bool probed1 = ((C >> (32 + lane)) & 1ull) && !((PRE >> (32 + lane)) & 1ull);
            const uint32_t moved = ballot(probed1 && c1 != pre.c[1]);
            const uint32_t limit = moved ? (uint32_t)(ffs(moved) - 1) : 32;
            cut = 32 + k1_commit_half(table, pre.h[1], p1, c1, (uint32_t)(C >> 32), (uint32_t)(PRE >> 32), limit, pre.h[1], unused);
        }
    }
    // ---- events of the accepted copies
    CS &= k1_below64(cut);
    const uint32_t ncopy = (uint32_t)popc((uint32_t)CS) + (uint32_t)popc((uint32_t)(CS >> 32));
    if (ncopy) {
        k1_wait_space(ring, head, ncopy);
        const uint32_t lo = (uint32_t)CS, hi = (uint32_t)(CS >> 32);
        const uint32_t below = (1u << lane) - 1u;
        if ((lo >> lane) & 1u) ring.ev[(head.head + popc(lo & below)) & (ring.size - 1)] = k1_event(p0, L0, p0 - pre.c[0]);
        if ((hi >> lane) & 1u) ring.ev[(head.head + popc(lo) + popc(hi & below)) & (ring.size - 1)] = k1_event(p1, L1, p1 - pre.c[1]);
        head.head += ncopy;
        if (head.head - head.published >= K1_PUBLISH) k1_publish(ring, head);
    }
    // ---- exit state
    if (cut < 64) {                                              // cut at a victim: restart the window there
#if defined(SB_EMU)
        if (lane == 0) g_k1_w64_stat[4]++;
#endif
        if (ncopy) {
            const uint32_t lastc = (CS >> 32) ? 63 - clz((uint32_t)(CS >> 32)) : 31 - clz((uint32_t)CS);
            const uint32_t e2 = lastc + k1_pick64(L0, L1, lastc);     // <= cut: a victim is never inside a copy
            if (e2 == cut) { st.s = w + cut; st.rematch = true; }
            else { st.s = w + cut; st.rematch = false; st.skip = 32 + (cut - e2 - 1); }
        } else {
            st.skip = st.rematch ? 32 + (cut - i0 - 1) : st.skip + (cut - i0);
            st.s = w + cut; st.rematch = false;
        }
        return true;
    }
    if (ncopy) {
        if (e_last >= 64) {
            st.s = w + e_last; st.rematch = true;
            if (e_last >= 65) {                                   // e-1 lies beyond this window
                if (nxt.w == w + 64 && e_last <= 128) {
                    if (st.s < s_limit) {                         // take its hash from the lane holding the prefetched words
                        const unsigned nsh = (unsigned)((uintptr_t)(win + nxt.w + lane) & 3u) * 8;
                        const uint32_t h0 = K1_HASH(funnel_r(nxt.a[0][0], nxt.a[0][1], nsh));
                        const uint32_t h1 = K1_HASH(funnel_r(nxt.a[1][0], nxt.a[1][1], nsh));
                        const uint32_t hsel = k1_pick64(h0, h1, e_last - 65);
                        syncwarp();
                        if (lane == 0) table[hsel] = (uint16_t)(st.s - 1);
                        syncwarp();
                    }
                } else {
                    k1_preinsert(win, table, shift, s_limit, st.s);
                }
            }
        } else {
            st.s = w + 64; st.rematch = false; st.skip = 32 + (63 - e_last);
        }
    } else {
        st.skip = st.rematch ? 32 + (63 - i0) : st.skip + (64 - i0);
        st.s = w + 64; st.rematch = false;
    }
    return true;
}

// Single parser warp over 64-position windows; same role as k1_parse_pipelined<1> (k1_compress.cuh).
template <bool GT>
SB_DEVICE void k1_parse64(const uint8_t* win, uint32_t n, uint16_t* table, const K1Ring& ring, uint32_t* ctrl) {
    const unsigned lane = lane_id();
    unsigned shift = 24;
    uint32_t tsize = 256;
    while (tsize < 16384 && tsize < n) { shift--; tsize *= 2; }   // src/compress.rs:491-497
    const uint32_t s_limit = n - 15;
    K1Prod prod;
    prod.head = ld_volatile(&ctrl[6]); prod.published = ld_volatile(&ctrl[7]); prod.tail_seen = 0;
    K1State st;
    st.s = 1; st.skip = 32; st.rematch = false;
    K1Seq64 seq;
    seq.w = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 5; k++) { seq.a[0][k] = 0; seq.a[1][k] = 0; }
    for (;;) {
        // A window starts where the parse stands (tools/sim_window_width.py: 60 bytes per step on text against 52 for
        // windows on 64-byte boundaries); the prefetch guesses that the next one starts at w + 64.
        const uint32_t w = K1_W64_ALIGNED_ON ? (st.s & ~63u) : st.s;
        // highest read of a step: aligned word of w+63 plus five words -> w + 82 < n
        const bool fast = w + 68 < s_limit && (st.rematch || st.skip < 64);
        bool finished;
        if (!fast && (st.rematch ? st.s >= s_limit : st.s + (st.skip >> 5) > s_limit)) finished = true;
        else {
            bool ok = false;
            if (fast) {
                K1Seq64 nxt = seq;
                if (nxt.w != w + 64 && w + 150 < n) nxt = k1_fetch_seq64(win, w + 64);   // next window's loads now
                const K1Pre64 pre = k1_eval64(win, table, shift, w, seq);
                seq = nxt;
#if defined(SB_EMU)
                const uint32_t s_before = st.s;
#endif
                ok = k1_finish64<GT>(win, n, table, shift, s_limit, st, ring, prod, pre, seq, w);
#if defined(SB_EMU)
                if (lane == 0) { if (ok) { g_k1_w64_stat[0]++; g_k1_w64_stat[1] += st.s - s_before; } else g_k1_w64_stat[2]++; }
#endif
            }
            if (!ok) finished = k1_serial(win, n, table, shift, s_limit, st, w + 64, ring, prod);
            else finished = false;
        }
        if (finished) {
            k1_push(ring, prod, k1_event(n, 0, 0));                // end marker -> trailing literal (:417-426)
            k1_publish(ring, prod);
            syncwarp();
            if (lane == 0) { ctrl[6] = prod.head; ctrl[7] = prod.published; }
            return;
        }
    }
}

#undef K1_HASH

}  // namespace sbk
