//! `stabilizer_stream::gpu` -- safe wrappers over the C ABI of libsspsd (include/sspsd.h) that
//! re-implement the public types of `src/psd.rs` on the B200.  SOURCE ONLY: never compiled here
//! (no Rust toolchain in the build container).  With this module the binaries keep their code:
//! `use stabilizer_stream::gpu::PsdCascade` instead of `stabilizer_stream::PsdCascade`.
use std::{ffi::CStr, ops::Range, os::raw::c_char, ptr};

pub use crate::{AvgOpts, Break, Detrend, MergeOpts};

#[repr(C)]
struct Config {
    n_fft: u32,
    window: i32,
    hbf: i32,
    device: i32,
    stream: *mut core::ffi::c_void,
    max_batch: u64,
    host_stage: u64,
    deep_defer: u64,
    flags: u32,
    _pad: u32,
}
#[repr(C)]
#[derive(Clone, Copy)]
struct CAvgOpts {
    limit: u32,
    count: u32,
}
#[repr(C)]
struct CMergeOpts {
    keep_overlap: u32,
    min_count: u32,
    keep_transition_band: u32,
}
#[repr(C)]
#[derive(Clone, Copy, Default)]
struct CBreak {
    start: u64,
    include: u32,
    count: u32,
    avg: u32,
    _pad: u32,
    bins_start: u64,
    bins_end: u64,
    fft_size: u64,
    decimation: u64,
    pending: u64,
    processed: u64,
}
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct CLoss {
    pub received: u64,
    pub dropped: u64,
    pub seq: u32,
    pub has_seq: u32,
}
#[repr(C)]
#[derive(Default)]
pub struct DecodeInfo {
    pub format: u32,
    pub n_traces: u32,
    pub samples_per_trace: u64,
    pub frames_ok: u64,
}
enum CCascade {}
enum CDecoder {}
enum CSource {}
enum CReceiver {}
enum CGroup {}

extern "C" {
    fn sspsd_last_error() -> *const c_char;
    fn sspsd_config_default(n_fft: u32, cfg: *mut Config) -> i32;
    fn sspsd_cascade_create(cfg: *const Config, out: *mut *mut CCascade) -> i32;
    fn sspsd_cascade_destroy(h: *mut CCascade);
    fn sspsd_cascade_clone(h: *mut CCascade, out: *mut *mut CCascade) -> i32;
    fn sspsd_cascade_process_f32(h: *mut CCascade, x: *const f32, n: usize, mem: i32) -> i32;
    fn sspsd_cascade_set_avg(h: *mut CCascade, avg: CAvgOpts) -> i32;
    fn sspsd_cascade_set_detrend(h: *mut CCascade, detrend: i32) -> i32;
    fn sspsd_cascade_rbw(h: *const CCascade, rbw: *mut f32) -> i32;
    fn sspsd_cascade_psd(h: *mut CCascade, opts: *const CMergeOpts, p: *mut f32, p_len: *mut usize,
                         b: *mut CBreak, b_len: *mut usize) -> i32;
    fn sspsd_decoder_create(device: i32, stream: *mut core::ffi::c_void, out: *mut *mut CDecoder) -> i32;
    fn sspsd_decoder_destroy(d: *mut CDecoder);
    fn sspsd_cascade_process_frames(d: *mut CDecoder, cascades: *const *mut CCascade, n: u32, frames: *const u8,
                                    n_frames: usize, frame_len: usize, frame_stride: usize, mem: i32,
                                    loss: *mut CLoss, info: *mut DecodeInfo) -> i32;
    fn sspsd_source_create(kind: i32, param: i64, seed: u64, device: i32, stream: *mut core::ffi::c_void,
                           out: *mut *mut CSource) -> i32;
    fn sspsd_source_destroy(s: *mut CSource);
    fn sspsd_cascade_process_source(h: *mut CCascade, s: *mut CSource, n: usize) -> i32;
    fn sspsd_receiver_create(ip: *const c_char, port: u16, slot_bytes: u32, n_slots: u32, flags: i32,
                             out: *mut *mut CReceiver) -> i32;
    fn sspsd_receiver_destroy(r: *mut CReceiver);
    fn sspsd_receiver_pump(r: *mut CReceiver, d: *mut CDecoder, cascades: *const *mut CCascade, n: u32, max_frames: u32,
                           timeout_ms: i32, loss: *mut CLoss, info: *mut DecodeInfo) -> i32;
    // multi-GPU partitioning inside the library (include/sspsd.h, "sspsd_group_*")
    fn sspsd_group_create(cfg: *const Config, devices: *const i32, n_devices: u32, shard_mode: i32, out: *mut *mut CGroup) -> i32;
    fn sspsd_group_destroy(g: *mut CGroup);
    fn sspsd_group_set_avg(g: *mut CGroup, avg: CAvgOpts) -> i32;
    fn sspsd_group_set_detrend(g: *mut CGroup, detrend: i32) -> i32;
    fn sspsd_group_process_f32(g: *mut CGroup, channel: u32, x: *const f32, n: usize, mem: i32) -> i32;
    fn sspsd_group_psd(g: *mut CGroup, channel: u32, opts: *const CMergeOpts, p: *mut f32, p_len: *mut usize,
                       b: *mut CBreak, b_len: *mut usize) -> i32;
    fn sspsd_group_time_plan(g: *mut CGroup, total: u64, n_local_stages: u32) -> i32;
    fn sspsd_group_time_process_all_f32(g: *mut CGroup, x: *const f32, n: usize) -> i32;
    fn sspsd_group_time_finish(g: *mut CGroup) -> i32;
}

fn check(status: i32) {
    // The reference's PSD API is infallible; a CUDA failure is as fatal as an allocation failure.
    if status != 0 {
        let msg = unsafe { CStr::from_ptr(sspsd_last_error()) }.to_string_lossy();
        panic!("sspsd error {status}: {msg}");
    }
}

/// Drop-in for `stabilizer_stream::PsdCascade<N>` (src/psd.rs:399-544).
pub struct PsdCascade<const N: usize> {
    h: *mut CCascade,
}
// moved into the receiver thread by the binaries (src/bin/psd.rs:170-176)
unsafe impl<const N: usize> Send for PsdCascade<N> {}

impl<const N: usize> Default for PsdCascade<N> {
    fn default() -> Self {
        let mut cfg = core::mem::MaybeUninit::<Config>::uninit();
        let mut h = ptr::null_mut();
        unsafe {
            check(sspsd_config_default(N as u32, cfg.as_mut_ptr()));
            check(sspsd_cascade_create(cfg.as_ptr(), &mut h));
        }
        Self { h }
    }
}

impl<const N: usize> Clone for PsdCascade<N> {
    fn clone(&self) -> Self {
        let mut h = ptr::null_mut();
        unsafe { check(sspsd_cascade_clone(self.h, &mut h)) };
        Self { h }
    }
}

impl<const N: usize> Drop for PsdCascade<N> {
    fn drop(&mut self) {
        unsafe { sspsd_cascade_destroy(self.h) }
    }
}

impl<const N: usize> PsdCascade<N> {
    pub fn rbw(&self) -> f32 {
        let mut r = 0.0;
        unsafe { check(sspsd_cascade_rbw(self.h, &mut r)) };
        r
    }
    pub fn set_avg(&mut self, avg: AvgOpts) {
        unsafe { check(sspsd_cascade_set_avg(self.h, CAvgOpts { limit: avg.limit, count: avg.count })) }
    }
    pub fn set_detrend(&mut self, d: Detrend) {
        // Detrend::Linear returns SSPSD_EUNIMPLEMENTED -> panics like unimplemented!() (src/psd.rs:110)
        unsafe { check(sspsd_cascade_set_detrend(self.h, d as i32)) }
    }
    pub fn process(&mut self, x: &[f32]) {
        unsafe { check(sspsd_cascade_process_f32(self.h, x.as_ptr(), x.len(), 0 /* SSPSD_MEM_HOST */)) }
    }
    pub fn psd(&self, opts: &MergeOpts) -> (Vec<f32>, Vec<Break>) {
        let o = CMergeOpts {
            keep_overlap: opts.keep_overlap as u32,
            min_count: opts.min_count,
            keep_transition_band: opts.keep_transition_band as u32,
        };
        let mut p = vec![0f32; 16 * (N / 2 + 1)];
        let mut b = vec![CBreak::default(); 16];
        let (mut pl, mut bl) = (p.len(), b.len());
        unsafe { check(sspsd_cascade_psd(self.h, &o, p.as_mut_ptr(), &mut pl, b.as_mut_ptr(), &mut bl)) };
        p.truncate(pl);
        let b = b[..bl]
            .iter()
            .map(|c| Break {
                start: c.start as usize,
                include: c.include != 0,
                count: c.count,
                avg: c.avg,
                bins: Range { start: c.bins_start as usize, end: c.bins_end as usize },
                fft_size: c.fft_size as usize,
                decimation: c.decimation as usize,
                pending: c.pending as usize,
                processed: c.processed as usize,
            })
            .collect();
        (p, b)
    }
}

/// Batched `Frame::from_bytes` + `Loss::update` + `traces()` feeding one cascade per trace
/// (src/source.rs:135-148 + src/bin/psd.rs:174-182 in one device pass).
pub struct FrameDecoder {
    d: *mut CDecoder,
}
unsafe impl Send for FrameDecoder {}

impl FrameDecoder {
    pub fn new(device: i32) -> Self {
        let mut d = ptr::null_mut();
        unsafe { check(sspsd_decoder_create(device, ptr::null_mut(), &mut d)) };
        Self { d }
    }
    /// `frames`: n equally sized frames back to back.  Errors map to `de::Error` by status code
    /// (5 InvalidHeader, 6 UnknownFormat, 7 PayloadSize; 8/9 are the reference's panics).
    pub fn process<const N: usize>(&mut self, cascades: &mut [PsdCascade<N>], frames: &[u8], frame_len: usize,
                                   loss: &mut CLoss) -> Result<DecodeInfo, i32> {
        let hs: Vec<*mut CCascade> = cascades.iter().map(|c| c.h).collect();
        let mut info = DecodeInfo::default();
        let st = unsafe {
            sspsd_cascade_process_frames(self.d, hs.as_ptr(), hs.len() as u32, frames.as_ptr(), frames.len() / frame_len,
                                         frame_len, frame_len, 0, loss, &mut info)
        };
        if st == 0 { Ok(info) } else { Err(st) }
    }
}

impl Drop for FrameDecoder {
    fn drop(&mut self) {
        unsafe { sspsd_decoder_destroy(self.d) }
    }
}

/// `Data::Noise` / `Data::Dsm` (src/source.rs:66-79, 104-134) generated in GPU memory: instead of
/// `source.get()` handing 4096 host samples to `PsdCascade::process`, `feed` produces and consumes
/// them on the device.
pub struct SyntheticSource {
    s: *mut CSource,
}
unsafe impl Send for SyntheticSource {}

impl SyntheticSource {
    /// `--noise e` (source.rs:40-42, seed as at source.rs:69)
    pub fn noise(exponent: i32) -> Self {
        let mut s = ptr::null_mut();
        unsafe { check(sspsd_source_create(0, exponent as i64, 0x7654321, 0, ptr::null_mut(), &mut s)) };
        Self { s }
    }
    /// `--dsm ftw` (source.rs:44-46)
    pub fn dsm(ftw: u32) -> Self {
        let mut s = ptr::null_mut();
        unsafe { check(sspsd_source_create(1, ftw as i64, 0, 0, ptr::null_mut(), &mut s)) };
        Self { s }
    }
    pub fn feed<const N: usize>(&mut self, cascade: &mut PsdCascade<N>, n: usize) {
        unsafe { check(sspsd_cascade_process_source(cascade.h, self.s, n)) }
    }
}

impl Drop for SyntheticSource {
    fn drop(&mut self) {
        unsafe { sspsd_source_destroy(self.s) }
    }
}

/// `Data::Udp` (src/source.rs:81-93, 159-165): datagrams are collected with recvmmsg into page-locked
/// slots and decoded per batch instead of `socket.read` + `Frame::from_bytes` per packet.
pub struct UdpReceiver {
    r: *mut CReceiver,
}
unsafe impl Send for UdpReceiver {}

impl UdpReceiver {
    /// `SourceOpts::{ip, port}` (source.rs:17-23)
    pub fn bind(ip: std::net::Ipv4Addr, port: u16) -> Self {
        let ip = std::ffi::CString::new(ip.to_string()).unwrap();
        let mut r = ptr::null_mut();
        unsafe { check(sspsd_receiver_create(ip.as_ptr(), port, 2048, 1024, 0, &mut r)) };
        Self { r }
    }
    /// One batch of datagrams -> decode -> loss -> one cascade per trace.  `Ok(info)` with
    /// `frames_ok == 0` after a second without traffic (the reference's read timeout, source.rs:83).
    pub fn pump<const N: usize>(&mut self, dec: &mut FrameDecoder, cascades: &mut [PsdCascade<N>], loss: &mut CLoss)
                                -> Result<DecodeInfo, i32> {
        let hs: Vec<*mut CCascade> = cascades.iter().map(|c| c.h).collect();
        let mut info = DecodeInfo::default();
        let st = unsafe { sspsd_receiver_pump(self.r, dec.d, hs.as_ptr(), hs.len() as u32, 1024, 1000, loss, &mut info) };
        if st == 0 { Ok(info) } else { Err(st) }
    }
}

impl Drop for UdpReceiver {
    fn drop(&mut self) {
        unsafe { sspsd_receiver_destroy(self.r) }
    }
}


fn to_break(c: &CBreak) -> Break {
    Break {
        start: c.start as usize,
        include: c.include != 0,
        count: c.count,
        avg: c.avg,
        bins: Range { start: c.bins_start as usize, end: c.bins_end as usize },
        fft_size: c.fft_size as usize,
        decimation: c.decimation as usize,
        pending: c.pending as usize,
        processed: c.processed as usize,
    }
}

/// The receiver loop's `Vec<(&str, PsdCascade<N>)>` (src/bin/psd.rs:170-183) spread over the GPUs of the box:
/// trace `c` lives on `devices[c % devices.len()]`; NCCL / peer loads stay inside the library.
pub struct Group<const N: usize> {
    g: *mut CGroup,
}
unsafe impl<const N: usize> Send for Group<N> {}

impl<const N: usize> Group<N> {
    fn new(devices: &[i32], mode: i32) -> Self {
        let mut cfg = std::mem::MaybeUninit::<Config>::uninit();
        let mut g = ptr::null_mut();
        unsafe {
            check(sspsd_config_default(N as u32, cfg.as_mut_ptr()));
            check(sspsd_group_create(cfg.as_ptr(), devices.as_ptr(), devices.len() as u32, mode, &mut g));
        }
        Self { g }
    }
    /// one cascade per trace (SSPSD_SHARD_CHANNELS)
    pub fn channels(devices: &[i32]) -> Self {
        Self::new(devices, 0)
    }
    /// one long capture cut into time chunks (SSPSD_SHARD_TIME)
    pub fn time_chunks(devices: &[i32]) -> Self {
        Self::new(devices, 1)
    }
    pub fn set_avg(&mut self, avg: AvgOpts) {
        unsafe { check(sspsd_group_set_avg(self.g, CAvgOpts { limit: avg.limit, count: avg.count })) }
    }
    pub fn set_detrend(&mut self, d: Detrend) {
        unsafe { check(sspsd_group_set_detrend(self.g, d as i32)) }
    }
    /// `dec[channel].process(&trace)`
    pub fn process(&mut self, channel: u32, x: &[f32]) {
        unsafe { check(sspsd_group_process_f32(self.g, channel, x.as_ptr(), x.len(), 0)) }
    }
    /// `dec[channel].psd(&merge_opts)`; for a time-chunked group (channel 0) after `time_finish`
    pub fn psd(&self, channel: u32, opts: &MergeOpts) -> (Vec<f32>, Vec<Break>) {
        let o = CMergeOpts { keep_overlap: opts.keep_overlap as u32, min_count: opts.min_count, keep_transition_band: opts.keep_transition_band as u32 };
        let mut p = vec![0f32; 16 * (N / 2 + 1)];
        let mut b = [CBreak::default(); 16];
        let (mut pl, mut bl) = (p.len(), b.len());
        unsafe { check(sspsd_group_psd(self.g, channel, &o, p.as_mut_ptr(), &mut pl, b.as_mut_ptr(), &mut bl)) };
        p.truncate(pl);
        (p, b[..bl].iter().map(to_break).collect())
    }
    pub fn time_plan(&mut self, total: u64, n_local_stages: u32) {
        unsafe { check(sspsd_group_time_plan(self.g, total, n_local_stages)) }
    }
    pub fn time_process_all(&mut self, x: &[f32]) {
        unsafe { check(sspsd_group_time_process_all_f32(self.g, x.as_ptr(), x.len())) }
    }
    pub fn time_finish(&mut self) {
        unsafe { check(sspsd_group_time_finish(self.g)) }
    }
}

impl<const N: usize> Drop for Group<N> {
    fn drop(&mut self) {
        unsafe { sspsd_group_destroy(self.g) }
    }
}
