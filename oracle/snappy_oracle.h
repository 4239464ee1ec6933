/*
 * snappy_oracle.h -- CPU restatement of rust-snappy (snap 1.1.1) used ONLY as a
 * test checker / CPU baseline. TEST INFRASTRUCTURE: nothing in the product
 * library (rust-snappy_b200/) may include, link or call this.
 *
 * Parity pin: reference golden vector test/tests.rs:199-205
 * (data/Mark.Twain-Tom.Sawyer.txt <-> .rawsnappy, 9871 bytes) and the decoder
 * KATs test/tests.rs:232-317, 345-466 (see tests/test_oracle.py).
 * The Rust crate itself cannot be built in this image (no rustc/cargo), so
 * there is no oracle/_ref; this restatement is what is pinned.
 */
#ifndef SNAPPY_ORACLE_H
#define SNAPPY_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Error variant indices follow the declaration order of `enum Error`
 * (reference src/error.rs:72-180). 0 = Ok. */
enum {
    ORC_OK = 0,
    ORC_TOO_BIG = 1,               /* a=given, b=max                       */
    ORC_BUFFER_TOO_SMALL = 2,      /* a=given, b=min                       */
    ORC_EMPTY = 3,
    ORC_HEADER = 4,
    ORC_HEADER_MISMATCH = 5,       /* a=expected_len, b=got_len            */
    ORC_LITERAL = 6,               /* a=len, b=src_len, c=dst_len          */
    ORC_COPY_READ = 7,             /* a=len, b=src_len                     */
    ORC_COPY_WRITE = 8,            /* a=len, b=dst_len                     */
    ORC_OFFSET = 9,                /* a=offset, b=dst_pos                  */
    ORC_STREAM_HEADER = 10,        /* a=byte                               */
    ORC_STREAM_HEADER_MISMATCH = 11, /* a=the 6 body bytes, little endian  */
    ORC_UNSUPPORTED_CHUNK_TYPE = 12, /* a=byte                             */
    ORC_UNSUPPORTED_CHUNK_LENGTH = 13, /* a=len, b=header(0/1)             */
    ORC_CHECKSUM = 14,             /* a=expected, b=got                    */
    ORC_IO_UNEXPECTED_EOF = 100    /* io::ErrorKind::UnexpectedEof from read_exact */
};

typedef struct {
    uint32_t code;
    uint32_t _pad;
    uint64_t a, b, c;
} orc_error;

/* src/compress.rs:42-53 */
size_t orc_max_compress_len(size_t input_len);
/* src/compress.rs:99-154 (Encoder::compress). Returns 0 on success. */
int orc_compress(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                 size_t *out_n, orc_error *err);
/* src/decompress.rs:30-35 */
int orc_decompress_len(const uint8_t *in, size_t n, size_t *out_len, orc_error *err);
/* src/decompress.rs:75-95 (Decoder::decompress) */
int orc_decompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                   size_t *out_n, orc_error *err);
/* src/crc32.rs:35-38 (masked) and the plain CRC-32C beneath it */
uint32_t orc_crc32c(const uint8_t *buf, size_t n);
uint32_t orc_crc32c_masked(const uint8_t *buf, size_t n);

/* src/frame.rs:62-104 compress_frame: writes the 8-byte chunk header+crc and
 * the chunk payload contiguously into out (cap >= 8 + 76490). */
int orc_compress_frame(const uint8_t *src, size_t n, uint8_t *out, size_t *out_n);

/* write::FrameEncoder semantics for `write_all(input); into_inner()` with ONE
 * write call (src/write.rs:123-192): stream identifier once, then one chunk per
 * <=65536-byte slice. Empty input writes nothing. cap >= orc_frame_max_len(n). */
size_t orc_frame_max_len(size_t n);
int orc_frame_encode(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_n);

/* read::FrameDecoder + read_to_end (src/read.rs:104-239) over an in-memory
 * stream. out must hold the decoded bytes (cap); on error returns the error the
 * reference would surface (snap::Error variant or UnexpectedEof) and *out_n =
 * bytes produced before the failure. If out==NULL only sizes are computed. */
int orc_frame_decode(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                     size_t *out_n, orc_error *err);

/* ---- CPU baseline drivers (multi-threaded, pthreads) -------------------- */
/* Compress `count` independent blocks: block i = text[off_i .. off_i+block_len)
 * with off_i = (first+i)*stride_mul % (text_len-block_len). Returns seconds of
 * wall time; *out_total = sum of compressed lengths. Each thread owns an
 * encoder (table) exactly like one rust Encoder per thread. */
double orc_bench_compress_mt(const uint8_t *text, size_t text_len, size_t block_len,
                             uint64_t first, uint64_t count, uint64_t stride_mul,
                             int threads, uint64_t *out_total);
/* Decompress `count` streams: stream i is streams[i % nstreams]; output into a
 * per-thread scratch buffer. Returns seconds; *out_total = decompressed bytes. */
double orc_fingerprint_blocks_mt(const uint8_t *text, size_t text_len, size_t block_len, uint64_t base, uint64_t step,
                                 uint64_t count, uint64_t stride_mul, int threads, uint32_t *out_lens, uint32_t *out_crcs);
double orc_bench_decompress_mt(const uint8_t *const *streams, const size_t *lens,
                               size_t nstreams, uint64_t count, int threads,
                               uint64_t *out_total);

#ifdef __cplusplus
}
#endif
#endif
