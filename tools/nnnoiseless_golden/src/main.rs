//! Runs nnnoiseless 0.5.2 -- through exactly the surface sleep3r/crispy uses (src-tauri/src/audio.rs:229
//! `DenoiseState::new()`, :268 `process_frame(&mut out[..], &in[..])`) -- over a raw f32 clip and writes what it
//! returns: the denoised frames and the per-frame VAD probability.
//!
//!   cargo run --release -- c1_input.f32 c1_output.f32 c1_vad.f32
//!
//! Input: little-endian f32 samples in 16-bit scale (the scale audio.rs:264 produces), a whole number of 480-sample
//! frames.  Output frame t is what process_frame wrote for input frame t (no first-frame drop: that is the caller's
//! business, audio.rs:275-278).
use nnnoiseless::{DenoiseState, FRAME_SIZE};
use std::fs;

fn read_f32(path: &str) -> Vec<f32> {
    let bytes = fs::read(path).unwrap_or_else(|e| panic!("cannot read {}: {}", path, e));
    assert!(bytes.len() % 4 == 0, "{} is not a whole number of f32", path);
    bytes.chunks_exact(4).map(|b| f32::from_le_bytes([b[0], b[1], b[2], b[3]])).collect()
}

fn write_f32(path: &str, data: &[f32]) {
    let mut bytes = Vec::with_capacity(data.len() * 4);
    for v in data {
        bytes.extend_from_slice(&v.to_le_bytes());
    }
    fs::write(path, bytes).unwrap_or_else(|e| panic!("cannot write {}: {}", path, e));
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    if args.len() != 4 {
        eprintln!("usage: {} <input.f32> <output.f32> <vad.f32>", args[0]);
        std::process::exit(2);
    }
    assert_eq!(FRAME_SIZE, 480);
    let input = read_f32(&args[1]);
    assert!(input.len() % FRAME_SIZE == 0, "input must be a whole number of {}-sample frames", FRAME_SIZE);
    let mut denoise = DenoiseState::new(); // Box<DenoiseState<'static>>, the crate's built-in model
    let mut output = vec![0f32; input.len()];
    let mut vad = Vec::with_capacity(input.len() / FRAME_SIZE);
    for (inp, out) in input.chunks_exact(FRAME_SIZE).zip(output.chunks_exact_mut(FRAME_SIZE)) {
        vad.push(denoise.process_frame(out, inp));
    }
    write_f32(&args[2], &output);
    write_f32(&args[3], &vad);
    eprintln!("nnnoiseless 0.5.2: {} frames denoised", vad.len());
}
