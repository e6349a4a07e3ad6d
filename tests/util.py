"""Shared helpers for the test-suite (test plumbing only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "libns_emu.so")
CSRC = os.path.join(ROOT, "crispy_b200", "csrc")

# tolerances stated by BASELINE.json north_star (16-bit scale: full scale = 32768)
TOL_MAX_ABS = 1e-3 * 32768.0
TOL_SNR_DB = 60.0
TOL_VAD = 1e-3


def build_emu() -> str:
    srcs = [os.path.join(EMU_DIR, "ns_emu.cpp"), os.path.join(CSRC, "ns_host.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("ns_pipe.cuh", "ns_common.h", "ns_simt.h", "ns_host.h")]
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off",
                               "-Wno-unknown-pragmas", "-o", EMU_LIB] + srcs)
    return EMU_LIB


_emu = None


def emu_lib():
    global _emu
    if _emu is None:
        L = C.CDLL(build_emu())
        L.ns_emu_process.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_longlong,
                                     C.c_longlong, C.c_int, C.c_uint, C.c_float, C.c_int]
        _emu = L
    return _emu


def emu_process(blob: bytes, x: np.ndarray, chunk: int = 16, flags: int = 0, volume: float = 1.0, state=None,
                out_dtype=np.float32, out_cols=None, app=None, out_frame_offset: int = 0):
    """Run the kernel body under the host SIMT emulation.  x: [n_streams, n_frames*480]."""
    L = emu_lib()
    x = np.ascontiguousarray(x)
    ns, n = x.shape
    nf = n // 480
    out = np.zeros((ns, out_cols if out_cols is not None else n), dtype=out_dtype)
    vad = np.zeros((ns, nf), np.float32)
    dbg = np.zeros((ns, nf, L.ns_emu_dbg_floats()), np.float32)
    if state is None:
        state = np.zeros((ns, L.ns_emu_state_floats()), np.float32)
    app_p, app_stride = None, 0
    if app is not None:
        app = np.ascontiguousarray(app, dtype=np.float32)
        app_p, app_stride = app.ctypes.data, app.shape[1]
    out_stride = out.shape[1] if not (flags & 8) else out.shape[1] // 2
    rc = L.ns_emu_process(blob, len(blob), x.ctypes.data, out.ctypes.data, vad.ctypes.data, app_p,
                          state.ctypes.data, dbg.ctypes.data, ns, nf, n, out_stride, app_stride, chunk, flags,
                          volume, out_frame_offset)
    assert rc == 0, f"ns_emu_process failed: {rc}"
    return out, vad, dbg, state


def snr_db(ref: np.ndarray, test: np.ndarray) -> float:
    ref = ref.astype(np.float64)
    err = test.astype(np.float64) - ref
    pe = float(np.mean(err ** 2))
    ps = float(np.mean(ref ** 2))
    if pe == 0.0:
        return 200.0
    return 10.0 * np.log10(max(ps, 1e-30) / pe)


def parity_report(ref_out, out, ref_vad, vad) -> dict:
    return {
        "max_abs": float(np.max(np.abs(out.astype(np.float64) - ref_out.astype(np.float64)))),
        "snr_db": snr_db(ref_out, out),
        "vad_max": float(np.max(np.abs(vad - ref_vad))) if vad.size else 0.0,
    }


def assert_parity(ref_out, out, ref_vad, vad, what=""):
    """north_star tolerances, in 16-bit scale: max abs <= 1e-3 FS, SNR >= 60 dB, VAD within 1e-3."""
    r = parity_report(ref_out, out, ref_vad, vad)
    assert r["max_abs"] <= TOL_MAX_ABS, f"{what}: max abs err {r['max_abs']} > {TOL_MAX_ABS}"
    assert r["snr_db"] >= TOL_SNR_DB, f"{what}: SNR {r['snr_db']} dB < {TOL_SNR_DB}"
    assert r["vad_max"] <= TOL_VAD, f"{what}: VAD err {r['vad_max']} > {TOL_VAD}"
    return r


def make_signal(n_streams: int, n_frames: int, seed: int = 0xC0FFEE, first_stream: int = 0) -> np.ndarray:
    """Synthetic speech+noise in 16-bit scale with an exact-silence stretch in stream 1 (if present)."""
    from crispy_b200.synth import synth_chunk
    x = synth_chunk(n_streams, n_frames * 480, seed=seed, first_stream=first_stream).numpy() * 32768.0
    x = x.astype(np.float32)
    if n_streams > 1 and n_frames >= 12:
        # digital silence exercises the E < 0.04 gate and its state rules.  After real audio the
        # high-pass rings for ~9 frames before the gate trips, so short tests put it at the start.
        x[1, : 4 * 480] = 0.0
        if n_frames >= 60:
            x[1, 20 * 480:45 * 480] = 0.0
    return x


ST_SYNTH = 1440  # ns_common.h kStSynth: two 480-float copies of synthesis_mem


def canonical_state(state: np.ndarray) -> np.ndarray:
    """The emulation keeps its chunk counter in the last word of stream 0's state block and
    synthesis_mem is double buffered on that counter's parity (ns_common.h kStSynth).  Return the
    state with the live synthesis_mem copy in slot 0, the stale copy and the counter cleared, so
    that states reached through different chunkings compare equal."""
    st = state.copy()
    sel = int(st.view(np.int32)[0, -1]) & 1
    live = st[:, ST_SYNTH + sel * 480: ST_SYNTH + (sel + 1) * 480].copy()
    st[:, ST_SYNTH: ST_SYNTH + 480] = live
    st[:, ST_SYNTH + 480: ST_SYNTH + 960] = 0.0
    st.view(np.int32)[0, -1] = 0
    return st


def adversarial_signals(n_frames: int) -> dict:
    """Inputs that push the pitch path to its corners (16-bit scale): DC, sparse impulses, loud and tiny noise, tones
    at and beyond the pitch range, clipping, a chirp, silence followed by a step."""
    n = n_frames * 480
    t = np.arange(n) / 48000.0
    rng = np.random.default_rng(1)
    sigs = {
        "dc": np.full(n, 5000.0),
        "impulses": np.where(np.arange(n) % 997 == 0, 30000.0, 0.0),
        "white_loud": rng.standard_normal(n) * 9000,
        "white_tiny": rng.standard_normal(n) * 0.3,
        "sine_60hz": 8000 * np.sin(2 * np.pi * 60 * t),
        "sine_800hz": 8000 * np.sin(2 * np.pi * 800 * t),
        "sine_62_5hz": 8000 * np.sin(2 * np.pi * 62.5 * t),
        "square_clip": np.clip(40000 * np.sin(2 * np.pi * 150 * t), -32768, 32767),
        "chirp": 9000 * np.sin(2 * np.pi * (80 + 4000 * t) * t),
        "zeros_then_step": np.concatenate([np.zeros(n // 2), np.full(n - n // 2, 12000.0)]),
    }
    return {k: v.astype(np.float32) for k, v in sigs.items()}
