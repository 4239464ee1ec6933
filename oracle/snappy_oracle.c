/*
 * snappy_oracle.c -- plain-C restatement of the rust-snappy raw/frame codec.
 * TEST INFRASTRUCTURE ONLY (checker + CPU baseline); see snappy_oracle.h.
 *
 * Every routine cites the reference lines whose observable behaviour it
 * restates (paths relative to the rust-snappy checkout).
 */
#define _GNU_SOURCE
#include "snappy_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define MAX_INPUT_SIZE 0xFFFFFFFFull /* src/lib.rs:93 */
#define MAX_BLOCK_SIZE 65536u        /* src/lib.rs:97 */
#define MAX_TABLE_SIZE 16384u        /* src/compress.rs:11 */
#define INPUT_MARGIN 15u             /* src/compress.rs:20 */
#define MIN_NON_LITERAL_BLOCK_SIZE 17u /* src/compress.rs:24 */
#define MAX_COMPRESS_BLOCK_SIZE 76490u /* src/frame.rs:12 */

static inline uint32_t ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t ld64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

static int fail(orc_error *e, uint32_t code, uint64_t a, uint64_t b, uint64_t c) {
    if (e) { e->code = code; e->_pad = 0; e->a = a; e->b = b; e->c = c; }
    return (int)code;
}
static void ok(orc_error *e) { if (e) { e->code = 0; e->_pad = 0; e->a = e->b = e->c = 0; } }

/* src/compress.rs:42-53 */
size_t orc_max_compress_len(size_t input_len) {
    uint64_t n = (uint64_t)input_len;
    if (n > MAX_INPUT_SIZE) return 0;
    uint64_t m = 32 + n + n / 6;
    return m > MAX_INPUT_SIZE ? 0 : (size_t)m;
}

/* src/bytes.rs:61-70 */
static size_t put_varint(uint8_t *dst, uint64_t v) {
    size_t i = 0;
    while (v >= 0x80) { dst[i++] = (uint8_t)v | 0x80; v >>= 7; }
    dst[i++] = (uint8_t)v;
    return i;
}

/* src/bytes.rs:73-90: returns header length, 0 when malformed. checked_shl
 * fails when shift >= 64 (it does NOT detect bits shifted out). */
static size_t get_varint(const uint8_t *p, size_t n, uint64_t *out) {
    uint64_t v = 0;
    unsigned shift = 0;
    for (size_t i = 0; i < n; i++) {
        uint8_t b = p[i];
        if (shift >= 64) return 0;
        if (b < 0x80) { *out = v | ((uint64_t)b << shift); return i + 1; }
        v |= (uint64_t)(b & 0x7F) << shift;
        shift += 7;
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Block encoder: src/compress.rs:195-474                                    */

static inline uint8_t *put_literal(uint8_t *d, const uint8_t *lit, size_t len) {
    /* src/compress.rs:433-474 (the 16-byte over-copy at :451 is an
     * optimisation with no effect on the first `len` bytes) */
    size_t m = len - 1;
    if (m <= 59) {
        *d++ = (uint8_t)(m << 2);
    } else if (m < 256) {
        *d++ = 60 << 2; *d++ = (uint8_t)m;
    } else {
        *d++ = 61 << 2; *d++ = (uint8_t)m; *d++ = (uint8_t)(m >> 8);
    }
    memcpy(d, lit, len);
    return d + len;
}

static inline uint8_t *put_copy2(uint8_t *d, size_t off, size_t len) {
    /* src/compress.rs:363-369 */
    *d++ = (uint8_t)(((len - 1) << 2) | 2);
    *d++ = (uint8_t)off; *d++ = (uint8_t)(off >> 8);
    return d;
}

static inline uint8_t *put_copy(uint8_t *d, size_t off, size_t len) {
    /* src/compress.rs:323-357 */
    while (len >= 68) { d = put_copy2(d, off, 64); len -= 64; }
    if (len > 64) { d = put_copy2(d, off, 60); len -= 60; }
    if (len <= 11 && off <= 2047) {
        *d++ = (uint8_t)(((off >> 8) << 5) | ((len - 4) << 2) | 1);
        *d++ = (uint8_t)off;
        return d;
    }
    return put_copy2(d, off, len);
}

/* One <=64KB block with n >= 17. `table` has room for 16384 entries. */
static uint8_t *encode_block(const uint8_t *src, size_t n, uint8_t *d, uint16_t *table) {
    /* table sizing: src/compress.rs:491-518 */
    unsigned shift = 24;
    size_t tsize = 256;
    while (tsize < MAX_TABLE_SIZE && tsize < n) { shift--; tsize *= 2; }
    memset(table, 0, tsize * sizeof(uint16_t));
#define HASH(x) (((uint32_t)(x) * 0x1E35A7BDu) >> shift) /* src/compress.rs:523-525 */

    const size_t s_limit = n - INPUT_MARGIN;
    size_t s = 1, next_emit = 0;
    uint32_t next_hash = HASH(ld32(src + 1));
    for (;;) {
        /* scan for a 4-byte match: src/compress.rs:204-245 */
        uint32_t skip = 32;
        size_t s_next = s, cand;
        for (;;) {
            s = s_next;
            uint32_t step = skip >> 5;
            s_next = s + step;
            skip += step;
            if (s_next > s_limit) goto finish;
            cand = table[next_hash];
            table[next_hash] = (uint16_t)s;
            next_hash = HASH(ld32(src + s_next));
            if (ld32(src + s) == ld32(src + cand)) break;
        }
        /* pending literal: src/compress.rs:250-257 */
        d = put_literal(d, src + next_emit, s - next_emit);
        /* copy run: src/compress.rs:258-315 */
        for (;;) {
            size_t base = s;
            s += 4;
            size_t c = cand + 4;
            /* extend up to the END OF THE BLOCK: src/compress.rs:378-412 */
            while (s + 8 <= n) {
                uint64_t z = ld64(src + s) ^ ld64(src + c);
                if (z == 0) { s += 8; c += 8; continue; }
                s += (size_t)__builtin_ctzll(z) >> 3;
                goto extended;
            }
            while (s < n && src[s] == src[c]) { s++; c++; }
        extended:
            d = put_copy(d, base - cand, s - base);
            next_emit = s;
            if (s >= s_limit) goto finish;
            uint64_t x = ld64(src + s - 1);
            table[HASH((uint32_t)x)] = (uint16_t)(s - 1);
            uint32_t h = HASH((uint32_t)(x >> 8));
            cand = table[h];
            table[h] = (uint16_t)s;
            if ((uint32_t)(x >> 8) != ld32(src + cand)) {
                next_hash = HASH((uint32_t)(x >> 16));
                s++;
                break;
            }
        }
    }
finish:
    /* src/compress.rs:417-426 */
    if (next_emit < n) d = put_literal(d, src + next_emit, n - next_emit);
    return d;
#undef HASH
}

static size_t encode_stream(const uint8_t *in, size_t n, uint8_t *out, uint16_t *table) {
    /* src/compress.rs:119-153 */
    if (n == 0) { out[0] = 0; return 1; }
    uint8_t *d = out + put_varint(out, (uint64_t)n);
    while (n) {
        size_t blk = n > MAX_BLOCK_SIZE ? MAX_BLOCK_SIZE : n;
        if (blk < MIN_NON_LITERAL_BLOCK_SIZE) d = put_literal(d, in, blk);
        else d = encode_block(in, blk, d, table);
        in += blk; n -= blk;
    }
    return (size_t)(d - out);
}

int orc_compress(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                 size_t *out_n, orc_error *err) {
    /* src/compress.rs:104-118 */
    size_t need = orc_max_compress_len(n);
    if (need == 0) return fail(err, ORC_TOO_BIG, (uint64_t)n, MAX_INPUT_SIZE, 0);
    if (cap < need) return fail(err, ORC_BUFFER_TOO_SMALL, (uint64_t)cap, (uint64_t)need, 0);
    uint16_t *table = (uint16_t *)malloc(MAX_TABLE_SIZE * sizeof(uint16_t));
    *out_n = encode_stream(in, n, out, table);
    free(table);
    ok(err);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Decoder: src/decompress.rs                                                */

int orc_decompress_len(const uint8_t *in, size_t n, size_t *out_len, orc_error *err) {
    /* src/decompress.rs:30-35, 362-374 */
    if (n == 0) { *out_len = 0; ok(err); return 0; }
    uint64_t v;
    size_t h = get_varint(in, n, &v);
    if (h == 0) return fail(err, ORC_HEADER, 0, 0, 0);
    if (v > MAX_INPUT_SIZE) return fail(err, ORC_TOO_BIG, v, MAX_INPUT_SIZE, 0);
    *out_len = (size_t)v;
    ok(err);
    return 0;
}

int orc_decompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                   size_t *out_n, orc_error *err) {
    /* src/decompress.rs:75-95 */
    if (n == 0) return fail(err, ORC_EMPTY, 0, 0, 0);
    uint64_t v;
    size_t hl = get_varint(in, n, &v);
    if (hl == 0) return fail(err, ORC_HEADER, 0, 0, 0);
    if (v > MAX_INPUT_SIZE) return fail(err, ORC_TOO_BIG, v, MAX_INPUT_SIZE, 0);
    if (v > (uint64_t)cap) return fail(err, ORC_BUFFER_TOO_SMALL, (uint64_t)cap, v, 0);

    const uint8_t *src = in + hl;
    const uint64_t sn = (uint64_t)(n - hl), dn = v;
    uint64_t s = 0, d = 0;
    /* element loop: src/decompress.rs:130-148. The reference's fast paths
     * (:170-186, :256-326) only over-copy inside dst and are output-equivalent
     * to the plain forms below on every successful decode. */
    while (s < sn) {
        uint8_t tag = src[s++];
        if ((tag & 3) == 0) {
            /* src/decompress.rs:161-228 */
            uint64_t len = (uint64_t)(tag >> 2) + 1;
            if (len >= 61) {
                if (s + 4 > sn) return fail(err, ORC_LITERAL, 4, sn - s, dn - d);
                unsigned nb = (unsigned)len - 60;
                uint32_t w = ld32(src + s);
                if (nb < 4) w &= (1u << (8 * nb)) - 1;
                len = (uint64_t)w + 1;
                s += nb;
            }
            if (sn - s < len || dn - d < len) return fail(err, ORC_LITERAL, len, sn - s, dn - d);
            /* same trick as the reference (:170-186): a fixed 16-byte move when there is room */
            if (len <= 16 && s + 16 <= sn && d + 16 <= dn) memcpy(out + d, src + s, 16);
            else memcpy(out + d, src + s, (size_t)len);
            s += len; d += len;
        } else {
            /* src/decompress.rs:233-343 and TagEntry::offset :433-474 */
            unsigned kind = tag & 3;
            unsigned nb = kind == 1 ? 1 : kind == 2 ? 2 : 4;
            uint64_t len = kind == 1 ? (uint64_t)(4 + ((tag >> 2) & 7)) : 1 + (uint64_t)(tag >> 2);
            uint64_t trailer;
            if (s + 4 <= sn) {
                uint32_t w = ld32(src + s);
                if (nb < 4) w &= (1u << (8 * nb)) - 1;
                trailer = w;
            } else if (nb == 1) {
                if (s >= sn) return fail(err, ORC_COPY_READ, 1, sn - s, 0);
                trailer = src[s];
            } else if (nb == 2) {
                if (s + 1 >= sn) return fail(err, ORC_COPY_READ, 2, sn - s, 0);
                trailer = (uint64_t)src[s] | ((uint64_t)src[s + 1] << 8);
            } else {
                return fail(err, ORC_COPY_READ, 4, sn - s, 0);
            }
            uint64_t off = trailer | (kind == 1 ? ((uint64_t)(tag >> 5) << 8) : 0);
            s += nb;
            if (off == 0 || d < off) return fail(err, ORC_OFFSET, off, d, 0);
            if (d + len > dn) return fail(err, ORC_COPY_WRITE, len, dn - d, 0);
            /* output-equivalent to the byte loop; 8-byte strides when source and
             * destination are at least 8 apart and the over-write stays inside dst
             * (the reference does the same with 16-byte strides, :256-326) */
            if (off >= 8 && d + len + 8 <= dn) {
                uint8_t *q = out + d;
                for (uint64_t i = 0; i < len; i += 8) memcpy(q + i, q + i - off, 8);
            } else {
                for (uint64_t i = 0; i < len; i++) out[d + i] = out[d + i - off];
            }
            d += len;
        }
    }
    if (d != dn) return fail(err, ORC_HEADER_MISMATCH, dn, d, 0);
    *out_n = (size_t)dn;
    ok(err);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* CRC-32C: src/crc32.rs, build.rs:69-124                                    */

static uint32_t crc_tab[8][256];
static pthread_once_t crc_once = PTHREAD_ONCE_INIT;
static void crc_init(void) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
        crc_tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int j = 1; j < 8; j++)
            crc_tab[j][i] = (crc_tab[j - 1][i] >> 8) ^ crc_tab[0][crc_tab[j - 1][i] & 0xFF];
}

uint32_t orc_crc32c(const uint8_t *buf, size_t n) {
    pthread_once(&crc_once, crc_init);
    uint32_t c = ~0u;
#if defined(__SSE4_2__)
    while (n >= 8) { c = (uint32_t)__builtin_ia32_crc32di(c, ld64(buf)); buf += 8; n -= 8; }
    while (n--) c = __builtin_ia32_crc32qi(c, *buf++);
#else
    while (n >= 8) {
        uint64_t w = ld64(buf) ^ c;
        c = crc_tab[7][w & 0xFF] ^ crc_tab[6][(w >> 8) & 0xFF] ^ crc_tab[5][(w >> 16) & 0xFF] ^
            crc_tab[4][(w >> 24) & 0xFF] ^ crc_tab[3][(w >> 32) & 0xFF] ^ crc_tab[2][(w >> 40) & 0xFF] ^
            crc_tab[1][(w >> 48) & 0xFF] ^ crc_tab[0][w >> 56];
        buf += 8; n -= 8;
    }
    while (n--) c = crc_tab[0][(c ^ *buf++) & 0xFF] ^ (c >> 8);
#endif
    return ~c;
}

/* bitwise reference used by the tests to cross-check the fast paths */
uint32_t orc_crc32c_bitwise(const uint8_t *buf, size_t n) {
    uint32_t c = ~0u;
    while (n--) {
        c ^= *buf++;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
    }
    return ~c;
}

uint32_t orc_crc32c_masked(const uint8_t *buf, size_t n) {
    /* src/crc32.rs:35-38 */
    uint32_t s = orc_crc32c(buf, n);
    return ((s >> 15) | (s << 17)) + 0xA282EAD8u;
}

/* ------------------------------------------------------------------------ */
/* Frame format: src/frame.rs, src/write.rs, src/read.rs                     */

static const uint8_t STREAM_IDENT[10] = {0xFF, 0x06, 0x00, 0x00, 's', 'N', 'a', 'P', 'p', 'Y'};

static size_t frame_chunk(const uint8_t *src, size_t n, uint8_t *out, uint16_t *table) {
    /* src/frame.rs:62-104 */
    uint32_t crc = orc_crc32c_masked(src, n);
    size_t clen = encode_stream(src, n, out + 8, table);
    size_t body;
    uint8_t type;
    if (clen >= n - n / 8) { type = 0x01; body = n; memcpy(out + 8, src, n); }
    else { type = 0x00; body = clen; }
    size_t chunk_len = 4 + body;
    out[0] = type;
    out[1] = (uint8_t)chunk_len; out[2] = (uint8_t)(chunk_len >> 8); out[3] = (uint8_t)(chunk_len >> 16);
    out[4] = (uint8_t)crc; out[5] = (uint8_t)(crc >> 8); out[6] = (uint8_t)(crc >> 16); out[7] = (uint8_t)(crc >> 24);
    return 8 + body;
}

int orc_compress_frame(const uint8_t *src, size_t n, uint8_t *out, size_t *out_n) {
    if (n > MAX_BLOCK_SIZE) return -1; /* assert at src/frame.rs:71 */
    uint16_t *table = (uint16_t *)malloc(MAX_TABLE_SIZE * sizeof(uint16_t));
    *out_n = frame_chunk(src, n, out, table);
    free(table);
    return 0;
}

size_t orc_frame_max_len(size_t n) {
    size_t chunks = (n + MAX_BLOCK_SIZE - 1) / MAX_BLOCK_SIZE;
    return 10 + chunks * (8 + MAX_COMPRESS_BLOCK_SIZE);
}

int orc_frame_encode(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_n) {
    /* src/write.rs:123-192 with a single write_all(in) followed by flush:
     * - n == 0: nothing is ever handed to Inner::write => empty output (:155-157)
     * - n <= 65536: staged in `src`, flushed as one chunk
     * - n  > 65536 with empty staging buffer: Inner::write(buf) splits into
     *   65536-byte slices, the last (short) slice is its own chunk (:132-135, :171-177) */
    if (n == 0) { *out_n = 0; return 0; }
    if (cap < orc_frame_max_len(n)) return -1;
    uint16_t *table = (uint16_t *)malloc(MAX_TABLE_SIZE * sizeof(uint16_t));
    uint8_t *tmp = (uint8_t *)malloc(8 + MAX_COMPRESS_BLOCK_SIZE);
    uint8_t *d = out;
    memcpy(d, STREAM_IDENT, 10); d += 10;
    while (n) {
        size_t blk = n > MAX_BLOCK_SIZE ? MAX_BLOCK_SIZE : n;
        size_t w = frame_chunk(in, blk, tmp, table);
        memcpy(d, tmp, w); d += w;
        in += blk; n -= blk;
    }
    free(tmp); free(table);
    *out_n = (size_t)(d - out);
    return 0;
}

int orc_frame_decode(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                     size_t *out_n, orc_error *err) {
    /* src/read.rs:104-239 driven by read_to_end over a slice reader. `src`
     * mirrors the decoder's persistent 76490-byte buffer because
     * decompress_len is applied to the WHOLE buffer (src/read.rs:216). */
    uint8_t *src = (uint8_t *)calloc(MAX_COMPRESS_BLOCK_SIZE, 1);
    uint8_t *dst = (uint8_t *)malloc(MAX_BLOCK_SIZE);
    size_t pos = 0, produced = 0;
    int seen_ident = 0, rc = 0;
    orc_error e; memset(&e, 0, sizeof e);
#define NEED(k) do { if (n - pos < (size_t)(k)) { rc = fail(&e, ORC_IO_UNEXPECTED_EOF, 0, 0, 0); goto done; } } while (0)
    for (;;) {
        if (pos == n) break; /* clean EOF: read_exact_eof -> Ok(false), src/read.rs:119-121 */
        NEED(4);
        memcpy(src, in + pos, 4); pos += 4;
        uint8_t ty = src[0];
        if (!seen_ident) {
            if (ty != 0xFF) { rc = fail(&e, ORC_STREAM_HEADER, ty, 0, 0); goto done; }
            seen_ident = 1;
        }
        uint64_t len = (uint64_t)src[1] | ((uint64_t)src[2] << 8) | ((uint64_t)src[3] << 16);
        if (len > MAX_COMPRESS_BLOCK_SIZE) { rc = fail(&e, ORC_UNSUPPORTED_CHUNK_LENGTH, len, 0, 0); goto done; }
        if (ty >= 0x02 && ty <= 0x7F) { rc = fail(&e, ORC_UNSUPPORTED_CHUNK_TYPE, ty, 0, 0); goto done; }
        if ((ty >= 0x80 && ty <= 0xFD) || ty == 0xFE) {
            NEED(len); memcpy(src, in + pos, len); pos += len;
        } else if (ty == 0xFF) {
            if (len != 6) { rc = fail(&e, ORC_UNSUPPORTED_CHUNK_LENGTH, len, 1, 0); goto done; }
            NEED(6); memcpy(src, in + pos, 6); pos += 6;
            if (memcmp(src, "sNaPpY", 6) != 0) {
                uint64_t a = 0; for (int i = 0; i < 6; i++) a |= (uint64_t)src[i] << (8 * i);
                rc = fail(&e, ORC_STREAM_HEADER_MISMATCH, a, 0, 0); goto done;
            }
        } else {
            if (len < 4) { rc = fail(&e, ORC_UNSUPPORTED_CHUNK_LENGTH, len, 0, 0); goto done; }
            NEED(4);
            uint32_t want = ld32(in + pos); pos += 4;
            size_t body = (size_t)len - 4, dn;
            if (ty == 0x01) {
                if (body > MAX_BLOCK_SIZE) { rc = fail(&e, ORC_UNSUPPORTED_CHUNK_LENGTH, body, 0, 0); goto done; }
                NEED(body); memcpy(dst, in + pos, body); pos += body;
                dn = body;
            } else {
                NEED(body); memcpy(src, in + pos, body); pos += body;
                rc = orc_decompress_len(src, MAX_COMPRESS_BLOCK_SIZE, &dn, &e);
                if (rc) goto done;
                if (dn > MAX_BLOCK_SIZE) { rc = fail(&e, ORC_UNSUPPORTED_CHUNK_LENGTH, dn, 0, 0); goto done; }
                size_t got;
                rc = orc_decompress(src, body, dst, dn, &got, &e);
                if (rc) goto done;
            }
            uint32_t have = orc_crc32c_masked(dst, dn);
            if (have != want) { rc = fail(&e, ORC_CHECKSUM, want, have, 0); goto done; }
            if (out) {
                if (produced + dn > cap) { rc = fail(&e, ORC_BUFFER_TOO_SMALL, cap, produced + dn, 0); goto done; }
                memcpy(out + produced, dst, dn);
            }
            produced += dn;
        }
    }
#undef NEED
done:
    free(src); free(dst);
    *out_n = produced;
    if (err) *err = e;
    return rc;
}

/* ------------------------------------------------------------------------ */
/* Multi-threaded CPU baseline drivers                                       */

static double now_s(void) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
    const uint8_t *text; size_t text_len, block_len; uint64_t first, count, mul;
    const uint8_t *const *streams; const size_t *lens; size_t nstreams;
    uint64_t total; int mode;
    uint64_t step, base; uint32_t *out_crcs, *out_lens;   /* mode 2: fingerprints of sampled blocks */
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    uint64_t total = 0;
    if (j->mode == 0) {
        uint16_t *table = (uint16_t *)malloc(MAX_TABLE_SIZE * sizeof(uint16_t));
        uint8_t *out = (uint8_t *)malloc(orc_max_compress_len(j->block_len));
        uint64_t span = (uint64_t)(j->text_len - j->block_len);
        for (uint64_t i = 0; i < j->count; i++) {
            uint64_t off = ((j->first + i) * j->mul) % span;
            total += encode_stream(j->text + off, j->block_len, out, table);
        }
        free(out); free(table);
    } else if (j->mode == 2) {
        /* sample k (k = first .. first+count) is block base + k*step: compressed length + masked CRC of the stream */
        uint16_t *table = (uint16_t *)malloc(MAX_TABLE_SIZE * sizeof(uint16_t));
        uint8_t *out = (uint8_t *)malloc(orc_max_compress_len(j->block_len));
        uint64_t span = (uint64_t)(j->text_len - j->block_len);
        for (uint64_t k = j->first; k < j->first + j->count; k++) {
            uint64_t off = ((j->base + k * j->step) * j->mul) % span;
            size_t n = encode_stream(j->text + off, j->block_len, out, table);
            j->out_lens[k] = (uint32_t)n;
            j->out_crcs[k] = orc_crc32c_masked(out, n);
            total += n;
        }
        free(out); free(table);
    } else {
        uint8_t *out = (uint8_t *)malloc(MAX_BLOCK_SIZE * 4);
        orc_error e;
        for (uint64_t i = 0; i < j->count; i++) {
            size_t k = (size_t)((j->first + i) % j->nstreams), got = 0;
            if (orc_decompress(j->streams[k], j->lens[k], out, MAX_BLOCK_SIZE * 4, &got, &e) == 0) total += got;
        }
        free(out);
    }
    j->total = total;
    return NULL;
}

static double run_mt(job_t proto, uint64_t count, int threads, uint64_t *out_total) {
    if (threads < 1) threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)threads);
    uint64_t per = count / (uint64_t)threads, rem = count % (uint64_t)threads, at = proto.first;
    double t0 = now_s();
    for (int t = 0; t < threads; t++) {
        jobs[t] = proto;
        jobs[t].first = at;
        jobs[t].count = per + ((uint64_t)t < rem ? 1 : 0);
        at += jobs[t].count;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    uint64_t total = 0;
    for (int t = 0; t < threads; t++) { pthread_join(th[t], NULL); total += jobs[t].total; }
    double t1 = now_s();
    free(th); free(jobs);
    if (out_total) *out_total = total;
    return t1 - t0;
}

double orc_bench_compress_mt(const uint8_t *text, size_t text_len, size_t block_len,
                             uint64_t first, uint64_t count, uint64_t stride_mul,
                             int threads, uint64_t *out_total) {
    job_t p; memset(&p, 0, sizeof p);
    p.text = text; p.text_len = text_len; p.block_len = block_len; p.first = first; p.mul = stride_mul; p.mode = 0;
    return run_mt(p, count, threads, out_total);
}

/* Parity fingerprints for bench.py: sample k in [0, count) is generator block base + k*step; its compressed
 * length and the masked CRC-32C of its compressed stream go to out_lens[k] / out_crcs[k]. Returns seconds. */
double orc_fingerprint_blocks_mt(const uint8_t *text, size_t text_len, size_t block_len, uint64_t base, uint64_t step,
                                 uint64_t count, uint64_t stride_mul, int threads, uint32_t *out_lens, uint32_t *out_crcs) {
    job_t p; memset(&p, 0, sizeof p);
    p.text = text; p.text_len = text_len; p.block_len = block_len; p.first = 0; p.mul = stride_mul; p.mode = 2;
    p.base = base; p.step = step; p.out_lens = out_lens; p.out_crcs = out_crcs;
    return run_mt(p, count, threads, NULL);
}

double orc_bench_decompress_mt(const uint8_t *const *streams, const size_t *lens,
                               size_t nstreams, uint64_t count, int threads,
                               uint64_t *out_total) {
    job_t p; memset(&p, 0, sizeof p);
    p.streams = streams; p.lens = lens; p.nstreams = nstreams; p.mode = 1;
    return run_mt(p, count, threads, out_total);
}
