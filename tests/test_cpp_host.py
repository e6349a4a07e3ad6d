"""The C++ host layer (include/crispy_ns.hpp): the compiled mirror of the reference's Rust operator interface
(audio.rs:73-134, :202-315; recording.rs:13-39, :78-127) above the C ABI.

tests/cpp/host_mirror_test.cpp restates the reference's own unit tests (audio.rs:1040-1096, recording.rs:406-520) and,
on a GPU, drives RnnNoiseProcessor / DenoiseState / BatchDenoiser / WavWriter; what it leaves in files is compared here
with the oracle (the oracle is test infrastructure: the program itself never links it)."""
import os
import subprocess

import numpy as np
import pytest

import crispy_b200 as cb
from crispy_b200 import build
from oracle import pyoracle as po
from tests.util import ROOT, make_signal, snr_db

SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
LIB_DIR = os.path.join(ROOT, "crispy_b200")


@pytest.fixture(scope="module", autouse=True)
def built():
    build.build()


def _compile(exe, extra=()):
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", *extra, "-I" + os.path.join(ROOT, "include"),
                           SRC, "-L" + LIB_DIR, "-lcrispy_ns", "-Wl,-rpath," + LIB_DIR, "-o", exe])


def _clips(tmp_path, n_frames=60):
    """a 48 kHz clip and a 44.1 kHz one, unit scale (what the capture callbacks hand to push_sample)"""
    x48 = (make_signal(1, n_frames, seed=0xC99)[0] / 32768.0).astype(np.float32)
    t = np.arange(n_frames * 441) / 44100.0
    rng = np.random.default_rng(0x441)
    x441 = (0.3 * np.sin(2 * np.pi * 180.0 * t) * (0.6 + 0.4 * np.sin(2 * np.pi * 3.1 * t)) + 0.12 * np.sin(2 * np.pi * 1270.0 * t)
            + 0.03 * rng.standard_normal(t.size)).astype(np.float32)
    x48.tofile(tmp_path / "clip48k.f32")
    x441.tofile(tmp_path / "clip441.f32")
    return x48, x441


@pytest.mark.parametrize("flags", [(), ("-march=native", "-ffp-contract=fast")], ids=["plain", "native-contract-fast"])
def test_cpp_mirror_restates_the_reference_unit_tests(tmp_path, flags):
    """LinearResampler and WavWriter behave as the reference's unit tests demand; the streaming interpolator gives the
    oracle's bits whatever the compiler is allowed to contract; without a device the constructors throw ENODEV."""
    exe = str(tmp_path / "host_mirror_test")
    _compile(exe, flags)
    _, x441 = _clips(tmp_path, 20)
    r = subprocess.run([exe, "cpu", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and "host_mirror_test: ok" in r.stdout, r.stdout + r.stderr
    got = np.fromfile(tmp_path / "out_linres.f32", np.float32)
    want = po.linear_resample(x441, 44100.0, 48000.0)
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))


def _next_sample_replay(pushed, input_rate, output_rate, pulls_per_frame):
    """audio.rs:297-314 over the frames push_sample returned, in the arithmetic of the Rust code (f64 position, f32 samples)"""
    buf, pos, out = [], 0.0, []
    step = float(np.float32(input_rate)) / float(np.float32(output_rate))
    for f in range(len(pushed) // 480):
        buf.extend(pushed[f * 480:(f + 1) * 480])
        for _ in range(pulls_per_frame):
            if len(buf) < 2:
                out.append(np.float32(0))
                continue
            silent = False
            while pos >= 1.0:
                buf.pop(0)
                pos -= 1.0
                if len(buf) < 2:
                    silent = True
                    break
            if silent:
                out.append(np.float32(0))
                continue
            s0, s1 = np.float32(buf[0]), np.float32(buf[1])
            frac = np.float32(pos)
            pos += step
            out.append(np.float32(s0 + np.float32(np.float32(s1 - s0) * frac)))
    return np.array(out, np.float32)


@pytest.mark.gpu
def test_cpp_operator_on_the_gpu_matches_the_oracle(oracle_model, tmp_path):
    exe = str(tmp_path / "host_mirror_test")
    _compile(exe)
    x48, x441 = _clips(tmp_path, 60)
    r = subprocess.run([exe, "gpu", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and "host_mirror_test: ok (gpu)" in r.stdout, r.stdout + r.stderr

    def load(name):
        return np.fromfile(tmp_path / name, np.float32)

    # (1) RnnNoiseProcessor at 48 kHz, volume 0.5: push_sample against the oracle's restatement of audio.rs:216-295
    pushed = load("out_push48k.f32")
    ref = po.processor_run(oracle_model, x48, 48000.0, 0.5)
    assert pushed.shape == ref.shape == (59 * 480,)
    assert np.max(np.abs(pushed - ref)) <= 1e-3 and snr_db(ref, pushed) >= 60
    # next_sample toward a 44.1 kHz device: the host arithmetic replayed on the very samples push_sample returned
    played = load("out_played441.f32")
    want = _next_sample_replay(pushed, 48000.0, 44100.0, 441)
    assert played.shape == want.shape and np.array_equal(played.view(np.uint32), want.view(np.uint32))
    assert np.count_nonzero(played) > 0.9 * played.size
    # (2) a 44.1 kHz microphone: LinearResampler in front, volume clamped to 1
    pushed441 = load("out_push441.f32")
    ref441 = po.processor_run(oracle_model, x441, 44100.0, 1.0)
    assert pushed441.shape == ref441.shape and pushed441.size >= 58 * 480
    assert np.max(np.abs(pushed441 - ref441)) <= 1e-3 and snr_db(ref441, pushed441) >= 60
    # (3) DenoiseState::process_frame in 16-bit scale + VAD
    o16, vad = load("out_frames16.f32"), load("out_vad.f32")
    r16, rvad = po.process_streams(oracle_model, (x48 * np.float32(32768.0))[None, :])
    assert np.max(np.abs(o16 - r16[0])) <= 1e-3 * 32768 and snr_db(r16[0], o16) >= 60
    assert np.max(np.abs(vad - rvad[0])) <= 1e-3
    # (4) app audio through resample_audio (bit-exact), then mic + app -> dual-mono PCM16 through WavWriter
    app48 = load("out_app48.f32")
    want_app = po.resample_audio(x441[: 60 * 441] * np.float32(0.25), 44100, 48000)
    assert np.array_equal(app48.view(np.uint32), want_app.view(np.uint32))
    pcm, sr = cb.wav_read_pcm16(str(tmp_path / "out_meeting.wav"))
    assert sr == 48000 and pcm.shape == (60 * 480, 2) and np.array_equal(pcm[:, 0], pcm[:, 1])
    refd, _ = po.process_streams(oracle_model, x48[None, :], unit_scale=True)
    want_mix = po.mix_dual_mono_i16(refd[0], app48).reshape(-1, 2)
    assert np.max(np.abs(pcm.astype(np.int32) - want_mix.astype(np.int32))) <= 34  # 1e-3 of full scale + the quantiser's step
