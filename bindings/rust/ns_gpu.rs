//! ns_gpu.rs -- drop-in for `nnnoiseless::{DenoiseState, FRAME_SIZE}` as sleep3r/crispy uses it
//! (src-tauri/src/audio.rs:4, :203, :229, :268), backed by libcrispy_ns.so (include/crispy_ns.h).
//!
//! Place in src-tauri/src/ and change audio.rs:4 to `use crate::ns_gpu::{DenoiseState, FRAME_SIZE};`.
//! NOT COMPILED IN THIS REPOSITORY: the build image has no Rust toolchain; the
//! same C entry points are exercised through ctypes by crispy_b200/_lib.py and the test-suite.
use std::os::raw::{c_char, c_float, c_int, c_void};

#[repr(C)]
pub struct CrispyNsModel {
    _p: [u8; 0],
}
#[repr(C)]
pub struct CrispyNsState {
    _p: [u8; 0],
}
#[repr(C)]
pub struct CrispyNsBatch {
    _p: [u8; 0],
}

pub const CRISPY_NS_IN_I16: u32 = 1 << 0;
pub const CRISPY_NS_OUT_I16: u32 = 1 << 1;
pub const CRISPY_NS_UNIT_SCALE: u32 = 1 << 2;
pub const CRISPY_NS_MIX_STEREO_I16: u32 = 1 << 3;
pub const CRISPY_NS_DROP_FIRST_FRAME: u32 = 1 << 8;

extern "C" {
    fn crispy_ns_last_error() -> *const c_char;
    fn crispy_ns_device_count() -> c_int;
    fn crispy_ns_create(model: *const CrispyNsModel, device: c_int, out: *mut *mut CrispyNsState) -> c_int;
    fn crispy_ns_process_frame(st: *mut CrispyNsState, out480: *mut c_float, in480: *const c_float, vad: *mut c_float) -> c_int;
    fn crispy_ns_reset(st: *mut CrispyNsState) -> c_int;
    fn crispy_ns_destroy(st: *mut CrispyNsState);
    fn crispy_ns_batch_create(model: *const CrispyNsModel, device: c_int, n_streams: c_int, out: *mut *mut CrispyNsBatch) -> c_int;
    fn crispy_ns_process_streams_host(
        b: *mut CrispyNsBatch, h_in: *const c_void, h_out: *mut c_void, h_vad: *mut c_float, h_app: *const c_float,
        n_frames: c_int, in_stride: i64, out_stride: i64, vad_stride: i64, app_stride: i64, flags: u32, volume: c_float,
    ) -> c_int;
    fn crispy_ns_batch_state_size(b: *const CrispyNsBatch) -> usize;
    fn crispy_ns_batch_save_state(b: *mut CrispyNsBatch, buf: *mut c_void, len: usize) -> c_int;
    fn crispy_ns_batch_load_state(b: *mut CrispyNsBatch, buf: *const c_void, len: usize) -> c_int;
    fn crispy_ns_batch_destroy(b: *mut CrispyNsBatch);
    fn crispy_ns_linear_resample_count(input_rate: c_float, output_rate: c_float, n_in: i64) -> i64;
    fn crispy_ns_sinc_resample_count(input_rate: c_int, output_rate: c_int, n_in: i64) -> i64;
    fn crispy_ns_resample_host(device: c_int, h_in: *const c_float, h_out: *mut c_float, n_streams: c_int, n_in: i64,
                               in_stride: i64, out_stride: i64, input_rate: c_int, output_rate: c_int, kind: c_int) -> c_int;
}

/// == nnnoiseless::FRAME_SIZE (audio.rs:4)
pub const FRAME_SIZE: usize = 480;

fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(crispy_ns_last_error()).to_string_lossy().into_owned() }
}

/// Same surface as `nnnoiseless::DenoiseState` as audio.rs uses it.
pub struct DenoiseState {
    h: *mut CrispyNsState,
}
// guarded by Mutex<NsState> exactly like today (audio.rs:693); the handle itself is not thread-safe
unsafe impl Send for DenoiseState {}

impl DenoiseState {
    /// audio.rs:229 `DenoiseState::new()`
    pub fn new() -> Box<DenoiseState> {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { crispy_ns_create(std::ptr::null(), 0, &mut h) };
        assert!(rc == 0, "crispy_ns_create failed: {}", last_error());
        Box::new(DenoiseState { h })
    }
    /// audio.rs:268 `process_frame(&mut out[..], &in[..])`: 480 f32 in 16-bit scale in and out; returns
    /// the VAD probability (the reference discards it).
    pub fn process_frame(&mut self, output: &mut [f32], input: &[f32]) -> f32 {
        assert_eq!(input.len(), FRAME_SIZE); // upstream asserts too
        assert_eq!(output.len(), FRAME_SIZE);
        let mut vad = 0f32;
        let rc = unsafe { crispy_ns_process_frame(self.h, output.as_mut_ptr(), input.as_ptr(), &mut vad) };
        assert!(rc == 0, "crispy_ns_process_frame failed: {}", last_error());
        vad
    }
    /// a fresh state, as the model switch at audio.rs:955-965 builds
    pub fn reset(&mut self) {
        let rc = unsafe { crispy_ns_reset(self.h) };
        assert!(rc == 0, "crispy_ns_reset failed: {}", last_error());
    }
}
impl Drop for DenoiseState {
    fn drop(&mut self) {
        unsafe { crispy_ns_destroy(self.h) }
    }
}

/// Many independent recordings at once (north_star `process_streams`): `n_streams` rows of
/// `n_frames * 480` unit-scale f32 samples, `stride` samples apart; arithmetic of
/// RnnNoiseProcessor::push_sample (audio.rs:261-278) fused into the kernels' load/store.
pub struct BatchDenoiser {
    h: *mut CrispyNsBatch,
    n_streams: usize,
}
unsafe impl Send for BatchDenoiser {}

impl BatchDenoiser {
    pub fn new(n_streams: usize, device: i32) -> Result<BatchDenoiser, String> {
        if unsafe { crispy_ns_device_count() } == 0 {
            return Err("no CUDA device: libcrispy_ns has no CPU fallback".into());
        }
        let mut h = std::ptr::null_mut();
        let rc = unsafe { crispy_ns_batch_create(std::ptr::null(), device, n_streams as c_int, &mut h) };
        if rc != 0 {
            return Err(last_error());
        }
        Ok(BatchDenoiser { h, n_streams })
    }
    /// State persists across calls, so a long recording can be fed in pieces.
    pub fn process_streams(&mut self, input: &[f32], output: &mut [f32], vad: Option<&mut [f32]>, n_frames: usize,
                           stride: usize, volume: f32, drop_first_frame: bool) -> Result<(), String> {
        assert!(input.len() >= (self.n_streams - 1) * stride + n_frames * FRAME_SIZE);
        assert!(output.len() >= (self.n_streams - 1) * stride + n_frames * FRAME_SIZE);
        let mut flags = CRISPY_NS_UNIT_SCALE;
        if drop_first_frame {
            flags |= CRISPY_NS_DROP_FIRST_FRAME;
        }
        let (vad_ptr, vad_stride) = match vad {
            Some(v) => {
                assert!(v.len() >= self.n_streams * n_frames);
                (v.as_mut_ptr(), n_frames as i64)
            }
            None => (std::ptr::null_mut(), 0),
        };
        let rc = unsafe {
            crispy_ns_process_streams_host(self.h, input.as_ptr() as *const c_void, output.as_mut_ptr() as *mut c_void, vad_ptr,
                                           std::ptr::null(), n_frames as c_int, stride as i64, stride as i64, vad_stride, 0,
                                           flags, volume)
        };
        if rc != 0 {
            return Err(last_error());
        }
        Ok(())
    }
    /// recorder path (commands/recording.rs:260-264 + recording.rs:108-110): mic denoised + app raw ->
    /// clamp -> interleaved dual-mono PCM16
    pub fn process_and_mix(&mut self, mic: &[f32], app: &[f32], out_pcm16: &mut [i16], n_frames: usize, stride: usize)
                           -> Result<(), String> {
        assert!(out_pcm16.len() >= 2 * ((self.n_streams - 1) * stride + n_frames * FRAME_SIZE));
        let rc = unsafe {
            crispy_ns_process_streams_host(self.h, mic.as_ptr() as *const c_void, out_pcm16.as_mut_ptr() as *mut c_void,
                                           std::ptr::null_mut(), app.as_ptr(), n_frames as c_int, stride as i64, stride as i64,
                                           0, stride as i64, CRISPY_NS_UNIT_SCALE | CRISPY_NS_MIX_STEREO_I16, 1.0)
        };
        if rc != 0 {
            return Err(last_error());
        }
        Ok(())
    }
    pub fn save_state(&mut self) -> Result<Vec<u8>, String> {
        let n = unsafe { crispy_ns_batch_state_size(self.h) };
        let mut buf = vec![0u8; n];
        let rc = unsafe { crispy_ns_batch_save_state(self.h, buf.as_mut_ptr() as *mut c_void, n) };
        if rc != 0 {
            return Err(last_error());
        }
        Ok(buf)
    }
    pub fn load_state(&mut self, buf: &[u8]) -> Result<(), String> {
        let rc = unsafe { crispy_ns_batch_load_state(self.h, buf.as_ptr() as *const c_void, buf.len()) };
        if rc != 0 {
            return Err(last_error());
        }
        Ok(())
    }
}
impl Drop for BatchDenoiser {
    fn drop(&mut self) {
        unsafe { crispy_ns_batch_destroy(self.h) }
    }
}

/// Front end for recordings that are not at 48 kHz (audio.rs:217-221 uses the linear interpolator; `sinc` selects the
/// windowed-sinc kernel, the rubato-style alternative).  `input` holds `n_streams` rows of `n_in` samples.
pub fn resample_to_48k(input: &[f32], n_streams: usize, n_in: usize, input_rate: u32, sinc: bool) -> Result<Vec<f32>, String> {
    let n_out = unsafe {
        if sinc { crispy_ns_sinc_resample_count(input_rate as c_int, 48000, n_in as i64) }
        else { crispy_ns_linear_resample_count(input_rate as c_float, 48000.0, n_in as i64) }
    } as usize;
    let mut out = vec![0f32; n_streams * n_out];
    let rc = unsafe {
        crispy_ns_resample_host(0, input.as_ptr(), out.as_mut_ptr(), n_streams as c_int, n_in as i64, n_in as i64,
                                n_out.max(1) as i64, input_rate as c_int, 48000, if sinc { 1 } else { 0 })
    };
    if rc == 0 { Ok(out) } else { Err(last_error()) }
}
