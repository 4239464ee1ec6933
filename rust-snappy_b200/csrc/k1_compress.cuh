// k1_compress.cuh -- K1: batched raw Snappy block encode, one <=64KB block per
// CTA, bit-exact with the reference encoder.
//
// Replaces reference src/compress.rs:195-317 (Block::compress), :323-369
// (emit_copy/emit_copy2), :378-412 (extend_match), :417-426 (done), :433-474
// (emit_literal), :491-526 (block_table + hash) and, per unit, the varint
// header + block loop of Encoder::compress (:119-153).
//
// The greedy parse is a serial dependency chain (every table insert depends on
// every earlier match decision), so bit-exactness forbids a "better" parallel
// match finder. The CTA is two warps around one block held in shared memory
// (64KB window + 16K-entry u16 hash table):
//
//  * PARSER warp. Looks at 32 consecutive positions at once. Every lane hashes
//    its position, reads the table as of the window start, fetches its
//    candidate and computes "would a probe here hit, and how long is the match".
//    From the hit bitmask each hit lane computes where the NEXT copy would start
//    (rematch hit at the copy end, else the first later hit of the scan), pointer
//    doubling from the window's entry state yields the copies the serial encoder
//    takes, and the inserted positions are "everything except copy interiors".
//    The inserts are committed, re-read, and if two inserted lanes collided on a
//    slot (the one case where a lane's candidate would have come from inside the
//    window) the window is undone and replayed by the serial path below.
//  * serial path: the reference's control flow executed by the warp (scan probes
//    32 at a time with match.any conflict resolution, 128-byte match extension).
//    Used for replays, for scan runs past 32 probes (stride > 1) and the block tail.
//  * EMITTER warp. Consumes the parser's (position, length, offset) copy events
//    from a shared-memory ring, 32 at a time: literal/copy tag sizes, a warp scan
//    for output offsets, tags and literal bytes written straight to HBM.
#pragma once
#include "common.cuh"

namespace sbk {

static const uint32_t K1_WIN_BYTES = 65536 + 64;             // window + slack for over-reads
static const uint32_t K1_TABLE_BYTES = 32768;                // 16K-entry u16 table
static const uint32_t K1_RING = 1024;                        // copy events in flight
static const uint32_t K1_SMEM_BYTES = K1_WIN_BYTES + K1_TABLE_BYTES + K1_RING * 8 + 64;
static const uint32_t K1_THREADS = 64;

SB_DEVICE uint32_t k1_rd32(const uint8_t* win, uint32_t p) {
    const uint32_t* w = (const uint32_t*)(win + (p & ~3u));
    return funnel_r(w[0], w[1], (p & 3u) * 8);
}

// ---------------------------------------------------------------- event ring
struct K1Ring {
    uint64_t* ev;        // K1_RING entries
    uint32_t* ctrl;      // [0]=head (produced), [1]=tail (consumed)
};
SB_DEVICE uint64_t k1_event(uint32_t pos, uint32_t len, uint32_t off) {
    return (uint64_t)pos | ((uint64_t)len << 17) | ((uint64_t)off << 34);
}
// producer side: `head` is the parser warp's private copy of ctrl[0]
SB_DEVICE void k1_wait_space(const K1Ring& r, uint32_t head, uint32_t need) {
    while (head + need - ld_volatile(&r.ctrl[1]) > K1_RING) spin();
}
SB_DEVICE void k1_publish(const K1Ring& r, uint32_t head) {
    threadfence_block();
    syncwarp();
    if (lane_id() == 0) st_volatile(&r.ctrl[0], head);
}
SB_DEVICE void k1_push(const K1Ring& r, uint32_t& head, uint64_t e) {
    k1_wait_space(r, head, 1);
    if (lane_id() == 0) r.ev[head % K1_RING] = e;
    head++;
    k1_publish(r, head);
}

// ------------------------------------------------------------- serial pieces
// match extension from (s, c) to the END OF THE BLOCK (src/compress.rs:378-412)
SB_DEVICE uint32_t k1_extend(const uint8_t* win, uint32_t n, uint32_t s, uint32_t c) {
    const unsigned lane = lane_id();
    for (;;) {
        const uint32_t p = s + 4 * lane;
        uint32_t m = 0;
        if (p < n) {
            const uint32_t avail = n - p;
            const uint32_t x = k1_rd32(win, p) ^ k1_rd32(win, c + 4 * lane);
            m = x ? (uint32_t)(ffs(x) - 1) >> 3 : 4;
            if (m > avail) m = avail;
        }
        const uint32_t stop = ballot(m < 4);
        if (!stop) { s += 128; c += 128; continue; }
        const unsigned f = ffs(stop) - 1;
        return s + 4 * f + shfl(m, f);
    }
}

struct K1State {
    uint32_t s;        // next event position
    uint32_t skip;     // scan state (src/compress.rs:204-211); meaningful when !rematch
    bool rematch;      // true: a copy just ended at s and s-1 is already inserted (:285-301 first half)
};

#define K1_HASH(x) (((uint32_t)(x) * 0x1E35A7BDu) >> shift)

// after a copy ends at e: `if s >= s_limit return` else insert e-1 (:275-295)
SB_DEVICE void k1_preinsert(const uint8_t* win, uint16_t* table, unsigned shift, uint32_t s_limit, uint32_t e) {
    if (e < s_limit) {
        const uint32_t h = K1_HASH(k1_rd32(win, e - 1));
        syncwarp();
        if (lane_id() == 0) table[h] = (uint16_t)(e - 1);
        syncwarp();
    }
}

// The reference's control flow, one event (or one 32-probe scan batch) at a time,
// until the parse position reaches `target` or the block is finished.
// Returns true when the block is finished.
SB_DEVICE bool k1_serial(const uint8_t* win, uint32_t n, uint16_t* table, unsigned shift, uint32_t s_limit,
                         K1State& st, uint32_t target, const K1Ring& ring, uint32_t& head) {
    const unsigned lane = lane_id();
    for (;;) {
        uint32_t cand;
        if (st.rematch) {
            if (st.s >= s_limit) return true;
            // probe at s (:296-313); s-1 was inserted when the copy ended
            const uint32_t cur = k1_rd32(win, st.s);
            const uint32_t h = K1_HASH(cur);
            cand = table[h];
            syncwarp();
            if (lane == 0) table[h] = (uint16_t)st.s;
            syncwarp();
            if (cur != k1_rd32(win, cand)) {
                st.s += 1; st.rematch = false; st.skip = 32;
                if (st.s >= target) return false;
                continue;
            }
        } else {
            // scan: 32 probes per step (:204-245)
            uint32_t pos = st.s, sk = st.skip;
            if (st.skip == 32) { pos = st.s + lane; sk = 32 + lane; }
            else { for (unsigned i = 0; i < lane; i++) { const uint32_t step = sk >> 5; pos += step; sk += step; } }
            const uint32_t step = sk >> 5;
            const bool valid = pos + step <= s_limit;          // probe happens only if s_next <= s_limit
            uint32_t cur = 0, h = 0xFFFFFFFFu - lane, c = 0;   // invalid lanes get unique pseudo-hashes
            if (valid) { cur = k1_rd32(win, pos); h = K1_HASH(cur); c = table[h]; }
            const uint32_t same = match_any(h);
            const uint32_t below = same & ((1u << lane) - 1u);
            const uint32_t prev_pos = shfl(pos, below ? 31 - clz(below) : 0);
            if (below) c = prev_pos;                            // an earlier probe of this batch inserted first
            const bool hit = valid && cur == k1_rd32(win, c);
            const uint32_t hm = ballot(hit), vm = ballot(valid);
            const unsigned fi = vm == 0xFFFFFFFFu ? 32 : ffs(~vm) - 1;
            const unsigned fh = hm ? ffs(hm) - 1 : 32;
            const unsigned ncommit = fh < fi ? fh + 1 : fi;    // lanes [0, ncommit) perform their insert
            if (lane < ncommit) {
                const uint32_t later = same & ~((2u << lane) - 1u) & (ncommit >= 32 ? 0xFFFFFFFFu : ((1u << ncommit) - 1u));
                if (!later) table[h] = (uint16_t)pos;           // last writer of a slot wins
            }
            syncwarp();
            if (fh >= fi) {
                if (fi < 32) return true;                       // ran past s_limit: block is finished
                st.s = shfl(pos + step, 31);
                st.skip = shfl(sk + step, 31);
                if (st.s >= target) return false;
                continue;
            }
            st.s = shfl(pos, fh);
            cand = shfl(c, fh);
        }
        // copy (:258-276)
        const uint32_t base = st.s;
        const uint32_t end = k1_extend(win, n, base + 4, cand + 4);
#ifdef SB_EMU_TRACE
        if (lane == 0) fprintf(stderr, "serial copy base=%u cand=%u len=%u\n", base, cand, end - base);
#endif
        k1_push(ring, head, k1_event(base, end - base, base - cand));
        k1_preinsert(win, table, shift, s_limit, end);
        st.s = end; st.rematch = true;
        if (st.s >= target) return false;
    }
}

// One 32-position window on the fast path. Returns false (state untouched, table
// restored) when the window must be replayed serially.
SB_DEVICE bool k1_window(const uint8_t* win, uint32_t n, uint16_t* table, unsigned shift, uint32_t s_limit,
                         K1State& st, const K1Ring& ring, uint32_t& head) {
    const unsigned lane = lane_id();
    const uint32_t w = st.s & ~31u, i0 = st.s - w, p = w + lane;
    // ---- speculative probe of every position against the table as of the window start
    const uint32_t* aw = (const uint32_t*)(win + (p & ~3u));
    const unsigned ash = (p & 3u) * 8;
    const uint32_t a0 = aw[0], a1 = aw[1], a2 = aw[2], a3 = aw[3];
    const uint32_t cur = funnel_r(a0, a1, ash);
    const uint32_t h = K1_HASH(cur);
    const uint32_t c = table[h];
    const uint32_t* bw = (const uint32_t*)(win + (c & ~3u));
    const unsigned bsh = (c & 3u) * 8;
    const uint32_t b0 = bw[0], b1 = bw[1], b2 = bw[2], b3 = bw[3];
    const bool eq = cur == funnel_r(b0, b1, bsh);
    uint32_t L = 4;                                             // match length, exact up to 11, 12 = "12 or more"
    {
        const uint32_t x4 = funnel_r(a1, a2, ash) ^ funnel_r(b1, b2, bsh);
        if (x4) L += (uint32_t)(ffs(x4) - 1) >> 3;
        else {
            const uint32_t x8 = funnel_r(a2, a3, ash) ^ funnel_r(b2, b3, bsh);
            L = 8 + (x8 ? (uint32_t)(ffs(x8) - 1) >> 3 : 4);
        }
    }
    const uint32_t E = ballot(eq);
    // first copy start from the entry state
    auto nextbit = [&](uint32_t x) -> uint32_t {
        if (x >= 32) return 32;
        const uint32_t m = E >> x;
        return m ? x + (uint32_t)(ffs(m) - 1) : 32;
    };
    uint32_t f;
    if (st.rematch) f = ((E >> i0) & 1u) ? i0 : nextbit(i0 + 1);
    else {
        f = nextbit(i0);
        const uint32_t probes = f < 32 ? f - i0 + 1 : 32 - i0;
        if (st.skip + probes > 64) return false;                 // the run leaves stride 1 inside this window
    }
    // ---- which hits are taken: pointer doubling over "next copy start"
    uint32_t longmask = ballot(eq && L == 12), CS = 0;
    for (;;) {
        const uint32_t e = lane + L;
        uint32_t nx = 64;
        if (eq && e < 32) { nx = ((E >> e) & 1u) ? e : nextbit(e + 1); if (nx >= 32) nx = 64; }
        uint32_t M = 1u << lane, T = nx;
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const uint32_t M2 = shfl(M, T & 31u), T2 = shfl(T, T & 31u);
            if (T < 32) { M |= M2; T = T2; }
        }
        CS = f < 32 ? shfl(M, f) : 0;
        const uint32_t unk = CS & longmask;
        if (!unk) break;
        // the first taken copy whose length is not exact yet: extend it cooperatively
        const unsigned j = ffs(unk) - 1;
        const uint32_t pj = w + j, cj = shfl(c, j);
        const uint32_t end = k1_extend(win, n, pj + 12, cj + 12);
        if (lane == j) L = end - pj;
        longmask &= ~(1u << j);
    }
    // ---- inserted positions = entry..31 minus copy interiors [q+1, e-2]
    const bool taken = (CS >> lane) & 1u;
    uint32_t interior = 0;
    if (taken && lane < 31) {
        const uint32_t lo = lane + 1, hi = lane + L - 2;         // L >= 4 -> hi >= lo
        const uint32_t upto = hi >= 31 ? 0xFFFFFFFFu : ((2u << hi) - 1u);
        interior = upto & ~((1u << lo) - 1u);
    }
    const uint32_t I = reduce_or(interior);
    const uint32_t C = (0xFFFFFFFFu << i0) & ~I;
    const bool ins = (C >> lane) & 1u;
#ifdef SB_EMU_TRACE
    if (lane == 0) fprintf(stderr, "win w=%u i0=%u rm=%d skip=%u E=%08x f=%u CS=%08x C=%08x\n", w, i0, (int)st.rematch, st.skip, E, f, CS, C);
#endif
    if (ins) table[h] = (uint16_t)p;
    syncwarp();
    const bool clash = ins && table[h] != (uint16_t)p;
    if (any(clash)) {                                            // two inserted lanes share a slot: undo, replay serially
        syncwarp();
        if (ins) table[h] = (uint16_t)c;
        syncwarp();
        return false;
    }
    // ---- publish the copies and leave the window
    const uint32_t ncopy = popc(CS);
    if (ncopy) {
        k1_wait_space(ring, head, ncopy);
        if (taken) ring.ev[(head + popc(CS & ((1u << lane) - 1u))) % K1_RING] = k1_event(p, L, p - c);
        head += ncopy;
        k1_publish(ring, head);
        const unsigned last = 31 - clz(CS);
        const uint32_t e_last = last + shfl(L, last);
        if (e_last >= 32) {
            st.s = w + e_last; st.rematch = true;
            if (e_last >= 33) k1_preinsert(win, table, shift, s_limit, st.s);   // e-1 lies beyond this window
        } else {
            st.s = w + 32; st.rematch = false; st.skip = 32 + (31 - e_last);
        }
    } else {
        st.skip = st.rematch ? 32 + (31 - i0) : st.skip + (32 - i0);
        st.s = w + 32; st.rematch = false;
    }
    return true;
}

// Parser warp: block already in shared memory, n >= 17.
SB_DEVICE void k1_parse_block(const uint8_t* win, uint32_t n, uint16_t* table, const K1Ring& ring, uint32_t& head) {
    unsigned shift = 24;
    uint32_t tsize = 256;
    while (tsize < 16384 && tsize < n) { shift--; tsize *= 2; }   // src/compress.rs:491-497
    const uint32_t s_limit = n - 15;
    K1State st;
    st.s = 1; st.skip = 32; st.rematch = false;
    for (;;) {
        if (st.rematch) { if (st.s >= s_limit) break; }
        else if (st.s + (st.skip >> 5) > s_limit) break;
        const uint32_t w = st.s & ~31u;
        if (w + 32 < s_limit && (st.rematch || st.skip < 64) &&
            k1_window(win, n, table, shift, s_limit, st, ring, head))
            continue;
        if (k1_serial(win, n, table, shift, s_limit, st, w + 32, ring, head)) break;
    }
}
#undef K1_HASH

// ------------------------------------------------------------------- emitter
// Consumes copy events until the end marker (len == 0, pos == n); returns bytes written.
SB_DEVICE uint32_t k1_emit_block(const uint8_t* win, uint8_t* out, uint32_t d, const K1Ring& ring, uint32_t& tail) {
    const unsigned lane = lane_id();
    uint32_t prev_end = 0;
    for (;;) {
        uint32_t avail;
        for (;;) {
            avail = ld_volatile(&ring.ctrl[0]) - tail;
            avail = shfl(avail, 0);
            if (avail) break;
            spin();
        }
        const uint32_t m = avail < 32 ? avail : 32;
        threadfence_block();
        uint64_t ev = 0;
        if (lane < m) ev = ring.ev[(tail + lane) % K1_RING];
        const uint32_t pos = (uint32_t)(ev & 0x1FFFFu), len = (uint32_t)((ev >> 17) & 0x1FFFFu), off = (uint32_t)(ev >> 34);
        const bool act = lane < m;
        const bool is_end = act && len == 0;
        // literal in front of every event: [end of previous copy, pos)
        uint32_t pe = shfl_up(pos + len, 1);
        if (lane == 0) pe = prev_end;
        const uint32_t lit = act ? pos - pe : 0;
        uint32_t lhdr = 0;
        if (lit) lhdr = lit <= 60 ? 1 : lit <= 256 ? 2 : 3;                         // src/compress.rs:436-463
        // copy tags (src/compress.rs:339-356)
        uint32_t rem = len, n64 = 0, n60 = 0, fin = 0;
        if (act && len) {
            if (rem >= 68) { n64 = (rem - 68) / 64 + 1; rem -= 64 * n64; }
            if (rem > 64) { n60 = 1; rem -= 60; }
            fin = (rem <= 11 && off <= 2047) ? 2 : 3;
        }
        const uint32_t size = lhdr + lit + 3 * (n64 + n60) + fin;
        uint32_t incl = size;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const uint32_t t = shfl_up(incl, k);
            if (lane >= (unsigned)k) incl += t;
        }
        uint8_t* o = out + d + (incl - size);
        if (lit) {
            const uint32_t mm = lit - 1;
            if (lhdr == 1) o[0] = (uint8_t)(mm << 2);
            else if (lhdr == 2) { o[0] = 60 << 2; o[1] = (uint8_t)mm; }
            else { o[0] = 61 << 2; o[1] = (uint8_t)mm; o[2] = (uint8_t)(mm >> 8); }
            o += lhdr;
            if (lit <= 16) for (uint32_t k = 0; k < lit; k++) o[k] = win[pe + k];
        }
        // long literals: whole warp, one at a time
        uint32_t big = ballot(lit > 16);
        while (big) {
            const unsigned j = ffs(big) - 1;
            big &= big - 1;
            const uint32_t jl = shfl(lit, j), jp = shfl(pe, j);
            const uint32_t jo = shfl((uint32_t)(o - out), j);
            warp_copy(out + jo, win + jp, jl);
        }
        if (act && len) {
            o += lit;
            for (uint32_t k = 0; k < n64; k++) { o[0] = (63 << 2) | 2; o[1] = (uint8_t)off; o[2] = (uint8_t)(off >> 8); o += 3; }
            if (n60) { o[0] = (59 << 2) | 2; o[1] = (uint8_t)off; o[2] = (uint8_t)(off >> 8); o += 3; }
            if (fin == 2) { o[0] = (uint8_t)(((off >> 8) << 5) | ((rem - 4) << 2) | 1); o[1] = (uint8_t)off; }
            else { o[0] = (uint8_t)(((rem - 1) << 2) | 2); o[1] = (uint8_t)off; o[2] = (uint8_t)(off >> 8); }
        }
        d += shfl(incl, 31);
        prev_end = shfl(pos + len, m - 1);
        tail += m;
        syncwarp();
        if (lane == 0) st_volatile(&ring.ctrl[1], tail);
        if (any(is_end)) return d;
    }
}

// Kernel body: CTA = parser warp + emitter warp, one unit (<= 65536 bytes) at a time.
// flags bit0: write the varint(length) header in front of the block body.
SB_DEVICE void k1_compress_body(const BatchDesc& b, uint32_t flags) {
    uint8_t* win = smem();
    uint16_t* table = (uint16_t*)(win + K1_WIN_BYTES);
    K1Ring ring;
    ring.ev = (uint64_t*)(win + K1_WIN_BYTES + K1_TABLE_BYTES);
    ring.ctrl = (uint32_t*)(win + K1_WIN_BYTES + K1_TABLE_BYTES + K1_RING * 8);
    const unsigned lane = lane_id(), wid = warp_id();
    uint32_t head = 0, tail = 0;      // parser's / emitter's private ring counters (never reset)
    if (thread_idx() == 0) { ring.ctrl[0] = 0; ring.ctrl[1] = 0; }
    for (uint32_t u = block_idx(); u < b.count; u += grid_dim()) {
        const uint8_t* in = unit_in(b, u);
        const uint32_t n = unit_in_len(b, u);
        uint8_t* out = unit_out(b, u);
        uint32_t d = 0;
        if (flags & 1u) {                                                  // varint header (:120-128)
            if (n == 0) { if (thread_idx() == 0) { out[0] = 0; b.out_lens[u] = 1; } continue; }
            uint32_t v = n;
            while (v >= 0x80) { if (thread_idx() == 0) out[d] = (uint8_t)v | 0x80; v >>= 7; d++; }
            if (thread_idx() == 0) out[d] = (uint8_t)v;
            d++;
        }
        if (n == 0) { if (thread_idx() == 0) b.out_lens[u] = d; continue; }
        syncthreads();                                                     // previous unit fully drained
        // stage the block: each warp copies one half (split on a 16-byte boundary), zero the table
        {
            const uint32_t half = ((n / 2) + 15) & ~15u;
            if (wid == 0) warp_copy(win, in, half < n ? half : n);
            else if (half < n) warp_copy(win + half, in + half, n - half);
            if (thread_idx() < 16) ((uint32_t*)(win + ((n + 3) & ~3u)))[thread_idx()] = 0;   // defined bytes for over-reads
            uint32_t tsize = 256;
            while (tsize < 16384 && tsize < n) tsize *= 2;
            for (uint32_t i = thread_idx(); i < tsize / 2; i += K1_THREADS) ((uint32_t*)table)[i] = 0;   // (:514-516)
        }
        syncthreads();
        if (wid == 0) {
            if (n >= 17) k1_parse_block(win, n, table, ring, head);        // (:140-150)
            k1_push(ring, head, k1_event(n, 0, 0));                        // end marker -> trailing literal (:417-426)
        } else {
            d = k1_emit_block(win, out, d, ring, tail);
            if (lane == 0) b.out_lens[u] = d;
        }
    }
}

}  // namespace sbk
