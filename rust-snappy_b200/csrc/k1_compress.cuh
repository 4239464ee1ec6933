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
// match finder. What the warp parallelises without changing the result:
//   * the 64KB input window and the u16 hash table live in shared memory;
//   * the scan loop (:207-245) evaluates the next 32 probe positions at once,
//     resolves same-hash conflicts inside the batch with match.any, takes the
//     first hit with a ballot and commits only the inserts the serial encoder
//     would have made before that hit;
//   * match extension (:378-412) compares 128 bytes per step;
//   * literal bytes are moved by all lanes.
#pragma once
#include "common.cuh"

namespace sbk {

static const uint32_t K1_WIN_BYTES = 65536 + 256;            // window + slack for over-reads
static const uint32_t K1_SMEM_BYTES = K1_WIN_BYTES + 32768;  // + 16K-entry u16 table

SB_DEVICE uint32_t k1_rd32(const uint8_t* win, uint32_t p) {
    const uint32_t* w = (const uint32_t*)(win + (p & ~3u));
    return funnel_r(w[0], w[1], (p & 3u) * 8);
}

struct K1Out {
    uint8_t* out;
    uint32_t d;
};

// tag bytes are written by lane 0, payload by the whole warp
SB_DEVICE void k1_emit_literal(K1Out& o, const uint8_t* win, uint32_t from, uint32_t len) {
    const unsigned lane = lane_id();
    const uint32_t m = len - 1;
    uint32_t h;
    if (m <= 59) { if (lane == 0) o.out[o.d] = (uint8_t)(m << 2); h = 1; }
    else if (m < 256) { if (lane == 0) { o.out[o.d] = 60 << 2; o.out[o.d + 1] = (uint8_t)m; } h = 2; }
    else { if (lane == 0) { o.out[o.d] = 61 << 2; o.out[o.d + 1] = (uint8_t)m; o.out[o.d + 2] = (uint8_t)(m >> 8); } h = 3; }
    o.d += h;
    warp_copy(o.out + o.d, win + from, len);
    o.d += len;
}

SB_DEVICE void k1_emit_copy(K1Out& o, uint32_t off, uint32_t len) {
    const bool w = lane_id() == 0;
    while (len >= 68) {
        if (w) { o.out[o.d] = (63 << 2) | 2; o.out[o.d + 1] = (uint8_t)off; o.out[o.d + 2] = (uint8_t)(off >> 8); }
        o.d += 3; len -= 64;
    }
    if (len > 64) {
        if (w) { o.out[o.d] = (59 << 2) | 2; o.out[o.d + 1] = (uint8_t)off; o.out[o.d + 2] = (uint8_t)(off >> 8); }
        o.d += 3; len -= 60;
    }
    if (len <= 11 && off <= 2047) {
        if (w) { o.out[o.d] = (uint8_t)(((off >> 8) << 5) | ((len - 4) << 2) | 1); o.out[o.d + 1] = (uint8_t)off; }
        o.d += 2;
    } else {
        if (w) { o.out[o.d] = (uint8_t)(((len - 1) << 2) | 2); o.out[o.d + 1] = (uint8_t)off; o.out[o.d + 2] = (uint8_t)(off >> 8); }
        o.d += 3;
    }
}

// Encode one block already resident in shared memory (win[0..n)), n >= 17.
SB_DEVICE void k1_encode_block(const uint8_t* win, uint32_t n, uint16_t* table, K1Out& o) {
    const unsigned lane = lane_id();
    unsigned shift = 24;
    uint32_t tsize = 256;
    while (tsize < 16384 && tsize < n) { shift--; tsize *= 2; }
    for (uint32_t i = lane; i < tsize / 2; i += 32) ((uint32_t*)table)[i] = 0;
    syncwarp();
#define K1_HASH(x) (((uint32_t)(x) * 0x1E35A7BDu) >> shift)

    const uint32_t s_limit = n - 15;
    uint32_t s = 1, next_emit = 0;
    for (;;) {
        // ---------------- scan: 32 probes per step (src/compress.rs:204-245)
        uint32_t skip = 32, cand = 0;
        bool found = false;
        for (;;) {
            uint32_t pos = s, sk = skip;
            if (skip == 32) { pos = s + lane; sk = 32 + lane; }
            else { for (unsigned i = 0; i < lane; i++) { const uint32_t st = sk >> 5; pos += st; sk += st; } }
            const uint32_t step = sk >> 5;
            const bool valid = pos + step <= s_limit;         // probe happens only if s_next <= s_limit
            uint32_t cur = 0, h = 0xFFFFFFFFu - lane, c = 0;  // invalid lanes get unique pseudo-hashes
            if (valid) { cur = k1_rd32(win, pos); h = K1_HASH(cur); c = table[h]; }
            const uint32_t same = match_any(h);
            const uint32_t below = same & ((1u << lane) - 1u);
            const uint32_t prev_pos = shfl(pos, below ? 31 - clz(below) : 0);
            if (below) c = prev_pos;                           // an earlier probe of this batch inserted first
            const bool hit = valid && cur == k1_rd32(win, c);
            const uint32_t hm = ballot(hit), vm = ballot(valid);
            const unsigned fi = vm == 0xFFFFFFFFu ? 32 : ffs(~vm) - 1;
            const unsigned fh = hm ? ffs(hm) - 1 : 32;
            const unsigned ncommit = fh < fi ? fh + 1 : fi;   // lanes [0, ncommit) perform their insert
            if (lane < ncommit) {
                const uint32_t later = same & ~((2u << lane) - 1u) & (ncommit >= 32 ? 0xFFFFFFFFu : ((1u << ncommit) - 1u));
                if (!later) table[h] = (uint16_t)pos;          // last writer of a slot wins
            }
#ifdef SB_EMU_TRACE
            if (lane==0) fprintf(stderr,"scan s=%u skip=%u hm=%08x vm=%08x fi=%u fh=%u\n", s, skip, hm, vm, fi, fh);
#endif
            syncwarp();
            if (fh < fi) { s = shfl(pos, fh); cand = shfl(c, fh); found = true; break; }
            if (fi < 32) break;                                // ran past s_limit: block is finished
            s = shfl(pos + step, 31);
            skip = shfl(sk + step, 31);
        }
        if (!found) break;
        // ---------------- pending literal (:250-257)
        k1_emit_literal(o, win, next_emit, s - next_emit);
        // ---------------- copy run (:258-315)
        for (;;) {
            const uint32_t base = s;
            s += 4;
            uint32_t c4 = cand + 4;
            for (;;) {                                         // extend to the END OF THE BLOCK (:380,:408)
                const uint32_t p = s + 4 * lane;
                uint32_t m = 0;
                if (p < n) {
                    const uint32_t avail = n - p;
                    const uint32_t x = k1_rd32(win, p) ^ k1_rd32(win, c4 + 4 * lane);
                    m = x ? (uint32_t)(ffs(x) - 1) >> 3 : 4;
                    if (m > avail) m = avail;
                }
                const uint32_t stop = ballot(m < 4);
                if (!stop) { s += 128; c4 += 128; continue; }
                const unsigned f = ffs(stop) - 1;
                s += 4 * f + shfl(m, f);
                break;
            }
#ifdef SB_EMU_TRACE
            if (lane==0) fprintf(stderr,"copy base=%u cand=%u len=%u\n", base, cand, s-base);
#endif
            k1_emit_copy(o, base - cand, s - base);
            next_emit = s;
            if (s >= s_limit) goto finish;
            // (:285-314) all lanes read, then lane 0 inserts s-1 and s (in that order)
            const uint32_t x0 = k1_rd32(win, s - 1), x1 = k1_rd32(win, s + 3);
            const uint32_t curw = funnel_r(x0, x1, 8);
            const uint32_t h0 = K1_HASH(x0), h = K1_HASH(curw);
            cand = h == h0 ? s - 1 : table[h];
            syncwarp();
            if (lane == 0) { table[h0] = (uint16_t)(s - 1); table[h] = (uint16_t)s; }
            syncwarp();
            if (curw != k1_rd32(win, cand)) { s += 1; break; }
        }
    }
finish:
    if (next_emit < n) k1_emit_literal(o, win, next_emit, n - next_emit);   // (:417-426)
#undef K1_HASH
}

// Kernel body: CTA = one warp = one unit (<= 65536 bytes) at a time.
// flags bit0: write the varint(length) header in front of the block body.
SB_DEVICE void k1_compress_body(const BatchDesc& b, uint32_t flags) {
    uint8_t* win = smem();
    uint16_t* table = (uint16_t*)(win + K1_WIN_BYTES);
    const unsigned lane = lane_id();
    for (uint32_t u = block_idx(); u < b.count; u += grid_dim()) {
        const uint8_t* in = unit_in(b, u);
        const uint32_t n = unit_in_len(b, u);
        K1Out o;
        o.out = unit_out(b, u);
        o.d = 0;
        if (flags & 1u) {
            if (n == 0) { if (lane == 0) { o.out[0] = 0; b.out_lens[u] = 1; } continue; }   // (:120-125)
            uint32_t v = n;
            while (v >= 0x80) { if (lane == 0) o.out[o.d] = (uint8_t)v | 0x80; v >>= 7; o.d++; }
            if (lane == 0) o.out[o.d] = (uint8_t)v;
            o.d++;
        }
        if (n > 0) {
#ifdef SB_EMU_TRACE
            if (lane==0) fprintf(stderr,"unit %u n=%u\n", u, n);
#endif
            syncwarp();
            warp_copy(win, in, n);
#ifdef SB_EMU_TRACE
            if (lane==0) fprintf(stderr,"copied\n");
#endif
            if (lane < 8) ((uint32_t*)(win + ((n + 3) & ~3u)))[lane] = 0;   // defined bytes for over-reads
            syncwarp();
            if (n < 17) k1_emit_literal(o, win, 0, n);                      // (:140-146)
            else k1_encode_block(win, n, table, o);
        }
        if (lane == 0) b.out_lens[u] = o.d;
        syncwarp();
    }
}

}  // namespace sbk
