// k4_frame.cuh -- K4: frame assembly around K1/K3 (and small utility kernels).
//
// Replaces reference src/frame.rs:62-104 (compress_frame: chunk type decision
// `compressed_len >= n - n/8`, 8-byte header = type, u24 length, masked CRC) and
// the chunk loop of src/write.rs:165-192 for a device-resident input; also the
// block concatenation of Encoder::compress for inputs above 64KB
// (src/compress.rs:128-153).
#pragma once
#include "common.cuh"

namespace sbk {

struct FramePlan {
    const uint8_t* in;        // uncompressed input (device)
    uint64_t n;               // total bytes
    const uint8_t* slots;     // K1 output slots, stride kSlotStride
    const uint32_t* clens;    // K1 output length per chunk
    const uint32_t* crcs;     // masked CRC per chunk (frame mode)
    uint32_t nchunks;
    uint32_t frame;           // 1: frame chunks with 8-byte headers; 0: raw block concatenation
    uint64_t base;            // bytes already in front of chunk 0 (stream identifier / varint)
    uint32_t* csize;          // out: bytes each chunk occupies in the final stream
    uint64_t* offs;           // out: offset of each chunk in the final stream; offs[nchunks] = total
    uint8_t* out;             // final stream
};

SB_DEVICE uint32_t k4_chunk_len(const FramePlan& p, uint32_t i) {
    const uint64_t at = (uint64_t)i * kMaxBlock;
    const uint64_t left = p.n - at;
    return left > kMaxBlock ? kMaxBlock : (uint32_t)left;
}

// thread per chunk: final size of the chunk
SB_DEVICE void k4_sizes_body(const FramePlan& p) {
    const uint64_t i = (uint64_t)block_idx() * block_dim() + thread_idx();
    if (i >= p.nchunks) return;
    const uint32_t n = k4_chunk_len(p, (uint32_t)i);
    const uint32_t c = p.clens[i];
    if (p.frame) p.csize[i] = 8 + ((c >= n - n / 8) ? n : c);     // src/frame.rs:85
    else p.csize[i] = c;
}

// single CTA: exclusive scan csize -> offs (64-bit), offs[nchunks] = total
SB_DEVICE void k4_scan_body(const FramePlan& p) {
    uint64_t* sh = (uint64_t*)smem();      // block_dim entries
    const unsigned t = thread_idx(), nt = block_dim();
    const uint64_t per = ((uint64_t)p.nchunks + nt - 1) / nt;
    const uint64_t lo = per * t < p.nchunks ? per * t : p.nchunks;
    const uint64_t hi = lo + per < p.nchunks ? lo + per : p.nchunks;
    uint64_t sum = 0;
    for (uint64_t i = lo; i < hi; i++) sum += p.csize[i];
    sh[t] = sum;
    syncthreads();
    if (t == 0) {
        uint64_t run = p.base;
        for (unsigned k = 0; k < nt; k++) { const uint64_t v = sh[k]; sh[k] = run; run += v; }
        p.offs[p.nchunks] = run;
    }
    syncthreads();
    uint64_t run = sh[t];
    for (uint64_t i = lo; i < hi; i++) { p.offs[i] = run; run += p.csize[i]; }
}

// warp per chunk: header + body into the final stream
SB_DEVICE void k4_gather_body(const FramePlan& p) {
    const unsigned wpb = block_dim() >> 5, lane = lane_id();
    const uint64_t nwarps = (uint64_t)grid_dim() * wpb;
    for (uint64_t u = (uint64_t)block_idx() * wpb + warp_id(); u < p.nchunks; u += nwarps) {
        const uint32_t i = (uint32_t)u;
        uint8_t* dst = p.out + p.offs[i];
        const uint8_t* slot = p.slots + (uint64_t)i * kSlotStride;
        if (p.frame) {
            const uint32_t n = k4_chunk_len(p, i), c = p.clens[i];
            const bool raw = c >= n - n / 8;
            const uint32_t body = raw ? n : c, clen = 4 + body, crc = p.crcs[i];
            if (lane < 8) {
                const uint64_t hdr = (uint64_t)(raw ? 1u : 0u) | ((uint64_t)clen << 8) | ((uint64_t)crc << 32);
                dst[lane] = (uint8_t)(hdr >> (8 * lane));                    // src/frame.rs:91-93
            }
            warp_copy(dst + 8, raw ? p.in + (uint64_t)i * kMaxBlock : slot, body);
        } else {
            warp_copy(dst, slot, p.clens[i]);
        }
    }
}

// K6: synthetic input -- unit i = text[off_i .. off_i+len), off_i = ((first+i)*mul) % (text_len-len)
struct GenPlan {
    const uint8_t* text; uint64_t text_len; uint8_t* out; uint64_t stride;
    uint32_t len; uint64_t first, count, mul;
};
SB_DEVICE void k6_generate_body(const GenPlan& g) {
    const unsigned wpb = block_dim() >> 5;
    const uint64_t nwarps = (uint64_t)grid_dim() * wpb;
    const uint64_t span = g.text_len - g.len;
    for (uint64_t u = (uint64_t)block_idx() * wpb + warp_id(); u < g.count; u += nwarps) {
        const uint64_t off = span ? ((g.first + u) * g.mul) % span : 0;
        warp_copy(g.out + u * g.stride, g.text + off, g.len);
    }
}

// warp per unit: plain copies (uncompressed frame chunks on the decode side)
SB_DEVICE void k5_copy_units_body(const BatchDesc& b) {
    const unsigned wpb = block_dim() >> 5;
    const uint64_t nwarps = (uint64_t)grid_dim() * wpb;
    for (uint64_t u = (uint64_t)block_idx() * wpb + warp_id(); u < b.count; u += nwarps)
        warp_copy(unit_out(b, (uint32_t)u), unit_in(b, (uint32_t)u), unit_in_len(b, (uint32_t)u));
}

}  // namespace sbk
