// snap.hpp -- C++ host-side mirror of the `snap` crate's public API over the C ABI
// of libsnapb200.so (include/snapb200.h). Header only.
//
//   snap::raw::{max_compress_len, decompress_len, Encoder, Decoder}   reference src/raw.rs:13-14
//   snap::write::FrameEncoder<W>                                     reference src/write.rs:34-161
//   snap::read::{FrameDecoder<R>, FrameEncoder<R>}                   reference src/read.rs:47-363
//   snap::Error                                                      reference src/error.rs:72-186
//
// Same names, argument meaning and error behaviour as the reference; all codec
// work happens in the CUDA kernels behind the C ABI (there is no CPU fallback:
// without a device every compute call throws Error{code = SB_E_NO_DEVICE}).
// W needs `void write_all(const uint8_t*, size_t)`; R needs `size_t read(uint8_t*, size_t)`
// returning 0 at end of input (the shapes of io::Write / io::Read).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/snapb200.h"

namespace snap {

struct Error : std::runtime_error {
    sb_error e;
    explicit Error(const sb_error& err) : std::runtime_error(describe(err)), e(err) {}
    uint32_t code() const { return e.code; }
    bool operator==(const Error& o) const { return e.code == o.e.code && e.a == o.e.a && e.b == o.e.b && e.c == o.e.c; }
    static std::string describe(const sb_error& e) {
        static const char* names[] = {"Ok", "TooBig", "BufferTooSmall", "Empty", "Header", "HeaderMismatch", "Literal",
                                      "CopyRead", "CopyWrite", "Offset", "StreamHeader", "StreamHeaderMismatch",
                                      "UnsupportedChunkType", "UnsupportedChunkLength", "Checksum"};
        std::string n = e.code <= 14 ? names[e.code] : e.code == SB_IO_UNEXPECTED_EOF ? "UnexpectedEof"
                        : e.code == SB_E_NO_DEVICE ? "NoDevice" : "LibraryError";
        return n + "{" + std::to_string(e.a) + "," + std::to_string(e.b) + "," + std::to_string(e.c) + "}";
    }
};

inline void check(int rc, const sb_error& e) { if (rc) throw Error(e); }

namespace raw {

inline size_t max_compress_len(size_t n) { return sb_max_compress_len(n); }          // src/compress.rs:42
inline size_t decompress_len(const uint8_t* in, size_t n) {                          // src/decompress.rs:30
    size_t out = 0; sb_error e;
    check(sb_decompress_len(in, n, &out, &e), e);
    return out;
}

class Encoder {                                                                       // src/compress.rs:67-170
  public:
    size_t compress(const uint8_t* in, size_t n, uint8_t* out, size_t cap) {
        size_t w = 0; sb_error e;
        check(sb_compress(in, n, out, cap, &w, &e), e);
        return w;
    }
    std::vector<uint8_t> compress_vec(const uint8_t* in, size_t n) {
        std::vector<uint8_t> buf(max_compress_len(n) ? max_compress_len(n) : 1);
        buf.resize(compress(in, n, buf.data(), max_compress_len(n)));
        return buf;
    }
};

class Decoder {                                                                       // src/decompress.rs:45-111
  public:
    size_t decompress(const uint8_t* in, size_t n, uint8_t* out, size_t cap) {
        size_t w = 0; sb_error e;
        check(sb_decompress(in, n, out, cap, &w, &e), e);
        return w;
    }
    std::vector<uint8_t> decompress_vec(const uint8_t* in, size_t n) {
        std::vector<uint8_t> buf(decompress_len(in, n));
        buf.resize(decompress(in, n, buf.data(), buf.size()));
        return buf;
    }
};

}  // namespace raw

namespace frame {
constexpr size_t MAX_BLOCK_SIZE = 1 << 16;                 // src/lib.rs:97
constexpr size_t MAX_COMPRESS_BLOCK_SIZE = 76490;          // src/frame.rs:12
static const uint8_t STREAM_IDENTIFIER[10] = {0xFF, 0x06, 0x00, 0x00, 's', 'N', 'a', 'P', 'p', 'Y'};   // src/frame.rs:18
inline std::vector<uint8_t> encode_chunks(const uint8_t* in, size_t n, bool ident) {
    std::vector<uint8_t> out(sb_frame_max_len(n));
    size_t w = 0; sb_error e;
    check(sb_frame_encode_ex(in, n, out.data(), out.size(), &w, ident ? 1 : 0, &e), e);
    out.resize(w);
    return out;
}
}  // namespace frame

namespace write {

template <class W>
class FrameEncoder {                                                                  // src/write.rs:34-161
  public:
    explicit FrameEncoder(W w) : w_(std::move(w)) { src_.reserve(frame::MAX_BLOCK_SIZE); }
    ~FrameEncoder() { if (!taken_) { try { flush(); } catch (...) {} } }             // Drop flushes, errors ignored (:112-120)
    size_t write(const uint8_t* buf, size_t n) {                                      // (:123-152)
        size_t total = 0;
        for (;;) {
            const size_t free_ = frame::MAX_BLOCK_SIZE - src_.size();
            size_t took;
            if (n <= free_) break;
            if (src_.empty()) took = inner_write(buf, n);
            else { src_.insert(src_.end(), buf, buf + free_); flush(); took = free_; }
            buf += took; n -= took; total += took;
        }
        src_.insert(src_.end(), buf, buf + n);
        return total + n;
    }
    void write_all(const uint8_t* buf, size_t n) { write(buf, n); }
    void flush() {                                                                    // (:154-161)
        if (src_.empty()) return;
        inner_write(src_.data(), src_.size());
        src_.clear();
    }
    W into_inner() { flush(); taken_ = true; return std::move(w_); }                  // (:91-96)
    W& get_ref() { return w_; }
    W& get_mut() { return w_; }

  private:
    size_t inner_write(const uint8_t* buf, size_t n) {                                // Inner::write (:165-192)
        if (!wrote_ident_) { wrote_ident_ = true; w_.write_all(frame::STREAM_IDENTIFIER, 10); }
        if (n) { auto c = frame::encode_chunks(buf, n, false); w_.write_all(c.data(), c.size()); }
        return n;
    }
    W w_;
    std::vector<uint8_t> src_;
    bool wrote_ident_ = false, taken_ = false;
};

}  // namespace write

namespace read {

// read::FrameDecoder over a reader; the whole compressed stream is pulled and
// decoded in one batched device call on first use (the chunk state machine of
// src/read.rs:104-239 runs inside sb_frame_decode), then served from memory.
template <class R>
class FrameDecoder {
  public:
    explicit FrameDecoder(R r) : r_(std::move(r)) {}
    size_t read(uint8_t* buf, size_t n) {
        if (!loaded_) load();
        const size_t k = n < out_.size() - at_ ? n : out_.size() - at_;
        memcpy(buf, out_.data() + at_, k);
        at_ += k;
        if (k == 0 && n && pending_) { pending_ = false; throw Error(err_); }         // error surfaces after the good bytes
        return k;
    }
    R& get_ref() { return r_; }
    R& get_mut() { return r_; }
    R into_inner() { return std::move(r_); }

  private:
    void load() {
        loaded_ = true;
        std::vector<uint8_t> in;
        uint8_t tmp[1 << 16];
        for (size_t k; (k = r_.read(tmp, sizeof tmp)) != 0;) in.insert(in.end(), tmp, tmp + k);
        size_t total = 0; sb_error e;
        check(sb_frame_decode(in.data(), in.size(), nullptr, 0, &total, &e), e);
        out_.resize(total ? total : 1);
        int rc = sb_frame_decode(in.data(), in.size(), out_.data(), total, &total, &e);
        out_.resize(total);
        if (rc) { pending_ = true; err_ = e; }
    }
    R r_;
    std::vector<uint8_t> out_;
    size_t at_ = 0;
    bool loaded_ = false, pending_ = false;
    sb_error err_{};
};

template <class R>
class FrameEncoder {                                                                  // src/read.rs:272-410
  public:
    explicit FrameEncoder(R r) : r_(std::move(r)), src_(frame::MAX_BLOCK_SIZE) {}
    size_t read(uint8_t* buf, size_t n) {
        if (at_ >= dst_.size()) {
            const size_t got = r_.read(src_.data(), src_.size());                     // ONE underlying read per chunk (:378-381)
            if (!got) return 0;
            dst_ = frame::encode_chunks(src_.data(), got, !wrote_ident_);
            wrote_ident_ = true;
            at_ = 0;
        }
        const size_t k = n < dst_.size() - at_ ? n : dst_.size() - at_;
        memcpy(buf, dst_.data() + at_, k);
        at_ += k;
        return k;
    }
    R& get_ref() { return r_; }
    R& get_mut() { return r_; }

  private:
    R r_;
    std::vector<uint8_t> src_, dst_;
    size_t at_ = 0;
    bool wrote_ident_ = false;
};

}  // namespace read
}  // namespace snap
