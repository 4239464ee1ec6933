//! `snap` API surface (raw::{Encoder, Decoder, max_compress_len, decompress_len},
//! write::FrameEncoder, read::{FrameDecoder, FrameEncoder}, Error) forwarding the
//! hot path to the B200 kernels through the C ABI of include/snapb200.h.
//! Host code stays in Rust; nothing here compresses on the CPU.
use std::io;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct SbError { pub code: u32, pub _pad: u32, pub a: u64, pub b: u64, pub c: u64 }

extern "C" {
    fn sb_max_compress_len(n: usize) -> usize;
    fn sb_compress(inp: *const u8, n: usize, out: *mut u8, cap: usize, out_n: *mut usize, e: *mut SbError) -> i32;
    fn sb_decompress_len(inp: *const u8, n: usize, out_len: *mut usize, e: *mut SbError) -> i32;
    fn sb_decompress(inp: *const u8, n: usize, out: *mut u8, cap: usize, out_n: *mut usize, e: *mut SbError) -> i32;
    fn sb_frame_max_len(n: usize) -> usize;
    fn sb_frame_encode_ex(inp: *const u8, n: usize, out: *mut u8, cap: usize, out_n: *mut usize, ident: i32, e: *mut SbError) -> i32;
    fn sb_frame_decode(inp: *const u8, n: usize, out: *mut u8, cap: usize, out_n: *mut usize, e: *mut SbError) -> i32;
}

/// Same variants and payload fields as the reference's `snap::Error`.
#[derive(Clone, Debug, PartialEq, Eq)]
pub enum Error {
    TooBig { given: u64, max: u64 }, BufferTooSmall { given: u64, min: u64 }, Empty, Header,
    HeaderMismatch { expected_len: u64, got_len: u64 }, Literal { len: u64, src_len: u64, dst_len: u64 },
    CopyRead { len: u64, src_len: u64 }, CopyWrite { len: u64, dst_len: u64 }, Offset { offset: u64, dst_pos: u64 },
    StreamHeader { byte: u8 }, StreamHeaderMismatch { bytes: Vec<u8> }, UnsupportedChunkType { byte: u8 },
    UnsupportedChunkLength { len: u64, header: bool }, Checksum { expected: u32, got: u32 },
    /// library-level failure (no device / CUDA error); never produced by the reference
    Device { code: u32, detail: u64 },
}
pub type Result<T> = std::result::Result<T, Error>;

impl std::fmt::Display for Error { fn fmt(&self, f: &mut std::fmt::Formatter) -> std::fmt::Result { write!(f, "{:?}", self) } }
impl std::error::Error for Error {}
impl From<Error> for io::Error { fn from(e: Error) -> io::Error { io::Error::new(io::ErrorKind::Other, e) } }

fn to_err(e: SbError) -> Error {
    match e.code {
        1 => Error::TooBig { given: e.a, max: e.b }, 2 => Error::BufferTooSmall { given: e.a, min: e.b },
        3 => Error::Empty, 4 => Error::Header, 5 => Error::HeaderMismatch { expected_len: e.a, got_len: e.b },
        6 => Error::Literal { len: e.a, src_len: e.b, dst_len: e.c }, 7 => Error::CopyRead { len: e.a, src_len: e.b },
        8 => Error::CopyWrite { len: e.a, dst_len: e.b }, 9 => Error::Offset { offset: e.a, dst_pos: e.b },
        10 => Error::StreamHeader { byte: e.a as u8 },
        11 => Error::StreamHeaderMismatch { bytes: e.a.to_le_bytes()[..6].to_vec() },
        12 => Error::UnsupportedChunkType { byte: e.a as u8 },
        13 => Error::UnsupportedChunkLength { len: e.a, header: e.b != 0 },
        14 => Error::Checksum { expected: e.a as u32, got: e.b as u32 },
        c => Error::Device { code: c, detail: e.a },
    }
}

pub mod raw {
    use super::*;
    pub fn max_compress_len(n: usize) -> usize { unsafe { sb_max_compress_len(n) } }
    pub fn decompress_len(input: &[u8]) -> Result<usize> {
        let (mut n, mut e) = (0usize, SbError::default());
        if unsafe { sb_decompress_len(input.as_ptr(), input.len(), &mut n, &mut e) } != 0 { return Err(to_err(e)); }
        Ok(n)
    }
    #[derive(Debug, Default)] pub struct Encoder { _p: () }
    impl Encoder {
        pub fn new() -> Encoder { Encoder { _p: () } }
        pub fn compress(&mut self, input: &[u8], output: &mut [u8]) -> Result<usize> {
            let (mut n, mut e) = (0usize, SbError::default());
            let rc = unsafe { sb_compress(input.as_ptr(), input.len(), output.as_mut_ptr(), output.len(), &mut n, &mut e) };
            if rc != 0 { Err(to_err(e)) } else { Ok(n) }
        }
        pub fn compress_vec(&mut self, input: &[u8]) -> Result<Vec<u8>> {
            let mut buf = vec![0; max_compress_len(input.len())];
            let n = self.compress(input, &mut buf)?;
            buf.truncate(n);
            Ok(buf)
        }
    }
    #[derive(Clone, Debug, Default)] pub struct Decoder { _p: () }
    impl Decoder {
        pub fn new() -> Decoder { Decoder { _p: () } }
        pub fn decompress(&mut self, input: &[u8], output: &mut [u8]) -> Result<usize> {
            let (mut n, mut e) = (0usize, SbError::default());
            let rc = unsafe { sb_decompress(input.as_ptr(), input.len(), output.as_mut_ptr(), output.len(), &mut n, &mut e) };
            if rc != 0 { Err(to_err(e)) } else { Ok(n) }
        }
        pub fn decompress_vec(&mut self, input: &[u8]) -> Result<Vec<u8>> {
            let mut buf = vec![0; decompress_len(input)?];
            let n = self.decompress(input, &mut buf)?;
            buf.truncate(n);
            Ok(buf)
        }
    }
}

const MAX_BLOCK_SIZE: usize = 1 << 16;
const STREAM_IDENTIFIER: &[u8] = b"\xFF\x06\x00\x00sNaPpY";

fn encode_chunks(buf: &[u8], ident: bool) -> Result<Vec<u8>> {
    let cap = unsafe { sb_frame_max_len(buf.len()) };
    let mut out = vec![0u8; cap];
    let (mut n, mut e) = (0usize, SbError::default());
    let rc = unsafe { sb_frame_encode_ex(buf.as_ptr(), buf.len(), out.as_mut_ptr(), cap, &mut n, ident as i32, &mut e) };
    if rc != 0 { return Err(to_err(e)); }
    out.truncate(n);
    Ok(out)
}

pub mod write {
    use super::*;
    /// Same staging rules as the reference (src/write.rs:123-161): they fix the chunk boundaries.
    pub struct FrameEncoder<W: io::Write> { w: Option<W>, src: Vec<u8>, wrote_stream_ident: bool }
    impl<W: io::Write> FrameEncoder<W> {
        pub fn new(wtr: W) -> Self { FrameEncoder { w: Some(wtr), src: Vec::with_capacity(MAX_BLOCK_SIZE), wrote_stream_ident: false } }
        pub fn get_ref(&self) -> &W { self.w.as_ref().unwrap() }
        pub fn get_mut(&mut self) -> &mut W { self.w.as_mut().unwrap() }
        pub fn into_inner(mut self) -> io::Result<W> { io::Write::flush(&mut self)?; Ok(self.w.take().unwrap()) }
        fn inner_write(&mut self, buf: &[u8]) -> io::Result<usize> {
            let w = self.w.as_mut().unwrap();
            if !self.wrote_stream_ident { self.wrote_stream_ident = true; w.write_all(STREAM_IDENTIFIER)?; }
            if !buf.is_empty() { w.write_all(&encode_chunks(buf, false)?)?; }
            Ok(buf.len())
        }
    }
    impl<W: io::Write> io::Write for FrameEncoder<W> {
        fn write(&mut self, mut buf: &[u8]) -> io::Result<usize> {
            let mut total = 0;
            loop {
                let free = MAX_BLOCK_SIZE - self.src.len();
                let n = if buf.len() <= free { break } else if self.src.is_empty() { self.inner_write(buf)? } else {
                    self.src.extend_from_slice(&buf[..free]); self.flush()?; free };
                buf = &buf[n..]; total += n;
            }
            self.src.extend_from_slice(buf);
            Ok(total + buf.len())
        }
        fn flush(&mut self) -> io::Result<()> {
            if self.src.is_empty() { return Ok(()); }
            let src = std::mem::take(&mut self.src);
            self.inner_write(&src)?;
            self.src = src; self.src.clear();
            Ok(())
        }
    }
    impl<W: io::Write> Drop for FrameEncoder<W> { fn drop(&mut self) { if self.w.is_some() { let _ = io::Write::flush(self); } } }
}

pub mod read {
    use super::*;
    /// Pulls the compressed stream, decodes every chunk in one batched device call, serves from memory.
    pub struct FrameDecoder<R: io::Read> { r: R, out: Vec<u8>, at: usize, loaded: bool, pending: Option<Error> }
    impl<R: io::Read> FrameDecoder<R> {
        pub fn new(rdr: R) -> Self { FrameDecoder { r: rdr, out: vec![], at: 0, loaded: false, pending: None } }
        pub fn get_ref(&self) -> &R { &self.r }
        pub fn get_mut(&mut self) -> &mut R { &mut self.r }
        pub fn into_inner(self) -> R { self.r }
    }
    impl<R: io::Read> io::Read for FrameDecoder<R> {
        fn read(&mut self, buf: &mut [u8]) -> io::Result<usize> {
            if !self.loaded {
                self.loaded = true;
                let mut inp = vec![]; self.r.read_to_end(&mut inp)?;
                let (mut n, mut e) = (0usize, SbError::default());
                if unsafe { sb_frame_decode(inp.as_ptr(), inp.len(), std::ptr::null_mut(), 0, &mut n, &mut e) } != 0 { return Err(to_err(e).into()); }
                self.out = vec![0; n.max(1)];
                let rc = unsafe { sb_frame_decode(inp.as_ptr(), inp.len(), self.out.as_mut_ptr(), n, &mut n, &mut e) };
                self.out.truncate(n);
                if rc != 0 { self.pending = Some(if e.code == 100 { Error::Device { code: 100, detail: 0 } } else { to_err(e) }); }
            }
            let k = buf.len().min(self.out.len() - self.at);
            buf[..k].copy_from_slice(&self.out[self.at..self.at + k]); self.at += k;
            if k == 0 && !buf.is_empty() {
                if let Some(e) = self.pending.take() {
                    return Err(match e { Error::Device { code: 100, .. } => io::Error::from(io::ErrorKind::UnexpectedEof), e => e.into() });
                }
            }
            Ok(k)
        }
    }
    /// One chunk per underlying read() of <=64KB (src/read.rs:368-409).
    pub struct FrameEncoder<R: io::Read> { r: R, src: Vec<u8>, dst: Vec<u8>, at: usize, wrote_stream_ident: bool }
    impl<R: io::Read> FrameEncoder<R> {
        pub fn new(rdr: R) -> Self { FrameEncoder { r: rdr, src: vec![0; MAX_BLOCK_SIZE], dst: vec![], at: 0, wrote_stream_ident: false } }
        pub fn get_ref(&self) -> &R { &self.r }
        pub fn get_mut(&mut self) -> &mut R { &mut self.r }
    }
    impl<R: io::Read> io::Read for FrameEncoder<R> {
        fn read(&mut self, buf: &mut [u8]) -> io::Result<usize> {
            if self.at >= self.dst.len() {
                let n = self.r.read(&mut self.src)?;
                if n == 0 { return Ok(0); }
                self.dst = encode_chunks(&self.src[..n], !self.wrote_stream_ident)?; self.wrote_stream_ident = true; self.at = 0;
            }
            let k = buf.len().min(self.dst.len() - self.at);
            buf[..k].copy_from_slice(&self.dst[self.at..self.at + k]); self.at += k;
            Ok(k)
        }
    }
}
