// host_mirror_test.cpp -- exercises rust-snappy_b200/host/snap.hpp like the
// reference's test/tests.rs helpers (press/depress/write_frame_press/...).
// argv[1] = a corpus file. Exit 0 = pass, 3 = no CUDA device (expected on CPU boxes).
#include <cstdio>
#include <vector>

#include "../../rust-snappy_b200/host/snap.hpp"

struct VecWriter { std::vector<uint8_t> v; void write_all(const uint8_t* p, size_t n) { v.insert(v.end(), p, p + n); } };
struct SliceReader {
    const uint8_t* p; size_t n, at = 0;
    size_t read(uint8_t* b, size_t k) { size_t t = k < n - at ? k : n - at; memcpy(b, p + at, t); at += t; return t; }
};
static std::vector<uint8_t> read_all(auto& r) {
    std::vector<uint8_t> out; uint8_t tmp[70000];
    for (size_t k; (k = r.read(tmp, sizeof tmp)) != 0;) out.insert(out.end(), tmp, tmp + k);
    return out;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    std::vector<uint8_t> d(1 << 22);
    d.resize(fread(d.data(), 1, d.size(), f));
    fclose(f);
    try {
        if (snap::raw::max_compress_len(65536) != 76490) return 1;
        auto c = snap::raw::Encoder().compress_vec(d.data(), d.size());                 // press
        auto back = snap::raw::Decoder().decompress_vec(c.data(), c.size());            // depress
        if (back != d) { puts("raw round trip mismatch"); return 1; }
        snap::write::FrameEncoder<VecWriter> w{VecWriter{}};                            // write_frame_press
        w.write_all(d.data(), d.size());
        auto framed = w.into_inner().v;
        snap::read::FrameEncoder<SliceReader> re{SliceReader{d.data(), d.size()}};      // read_frame_press
        if (read_all(re) != framed) { puts("read/write frame encoders differ"); return 1; }
        snap::read::FrameDecoder<SliceReader> rd{SliceReader{framed.data(), framed.size()}};   // read_frame_depress
        if (read_all(rd) != d) { puts("frame round trip mismatch"); return 1; }
        const uint8_t bad[] = {0x05, 0x00, 'a'};                                        // err_header_mismatch KAT
        try { snap::raw::Decoder().decompress_vec(bad, 3); return 1; }
        catch (const snap::Error& e) { if (e.code() != SB_HEADER_MISMATCH || e.e.a != 5 || e.e.b != 1) return 1; }
        printf("host mirror ok: %zu -> %zu raw, %zu framed\n", d.size(), c.size(), framed.size());
        return 0;
    } catch (const snap::Error& e) {
        if (e.code() == SB_E_NO_DEVICE) { puts("no CUDA device: compute calls refused (no CPU fallback)"); return 3; }
        printf("unexpected error %s\n", e.what());
        return 1;
    }
}
