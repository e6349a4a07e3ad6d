"""GPU: parity against the oracle at the BASELINE.json configurations' real lengths (the CUDA path through the C ABI).

  configs[1]  the 1,024-stream x 60 s batch: 16 of its streams over all 6,000 frames
  configs[3]  dual-source meetings, 10 min: mic PCM16 + app f32 -> dual-mono stereo PCM16
  configs[4]  60-min streams fed as sixty 60 s calls with every DenoiseState carried
  adversarial inputs (DC, impulses, tones at the pitch range's edges, clipping ...)
Each test appends its figures (decision flips, max abs error, SNR, VAD error) to the parity report
($CRISPY_PARITY_REPORT, default gpurun_out/r2_parity.json; profiles/r2_parity.json is a committed copy).
Tolerances are BASELINE.json north_star's: max abs <= 1e-3 of full scale, SNR >= 60 dB, VAD within 1e-3.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import crispy_b200 as cb  # noqa: E402
from crispy_b200.shard import stream_block  # noqa: E402
from crispy_b200.synth import synth_chunk  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tests.util import BRANCH_EPS, adversarial_signals, long_run_parity, parity_report, snr_db  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.environ.get("CRISPY_PARITY_REPORT", os.path.join(ROOT, "gpurun_out", "r2_parity.json"))
CORES = os.cpu_count() or 1


def report(key: str, value: dict) -> None:
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        data = json.load(open(REPORT)) if os.path.exists(REPORT) else {}
        data[key] = value
        json.dump(data, open(REPORT, "w"), indent=1, sort_keys=True)
    except OSError:
        pass
    print(key, json.dumps(value))


@pytest.fixture(scope="module")
def model():
    return cb.Model.synthetic(0)


def bits_differ(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """float32 arrays compared for identity, with NaN equal to NaN: on a signal decaying into the denormal range
    RNNoise's LPC divides by a denormal error once and the frame's pitch gain is NaN (silent frame, no audible effect)
    -- in the oracle and in the kernels alike; their NaN payloads need not match."""
    return (a != b) & ~(np.isnan(a) & np.isnan(b))


def synth_device(n_streams: int, n_frames: int, first_stream: int = 0, start_frame: int = 0):
    return torch.cat([synth_chunk(n_streams, min(100, n_frames - f) * 480, first_stream=first_stream,
                                  start_sample=(start_frame + f) * 480, device="cuda") for f in range(0, n_frames, 100)], 1)


def test_c2_sixteen_streams_of_the_full_batch_over_60s(oracle_model, model):
    """configs[1] at full size on the GPU (1,024 streams x 6,000 frames); 16 of the streams, spread over the batch
    and including the muted ones (id % 16 == 3), against the oracle over all 6,000 frames: GRU / lastg drift over a
    BASELINE-length recording would show here."""
    n_streams, n_frames = 1024, 6000
    x = synth_device(n_streams, n_frames)
    den = cb.BatchDenoiser(n_streams, model)
    out, vad = den.process_streams(x, unit_scale=True)
    ids = [0, 3, 19, 64, 127, 200, 255, 256, 333, 511, 512, 640, 777, 900, 1003, 1023]
    xs = x[ids].cpu().numpy()
    ref, rvad, rpi, rpg, rsil, rmargin = po.process_streams_trace(oracle_model, xs, unit_scale=True, n_threads=CORES,
                                                                  native=True, margin=True)
    got, gv = out[ids].cpu().numpy(), vad[ids].cpu().numpy()
    # decisions of the same 16 streams from a second, 16-stream batch with taps (bit-identical samples, so the same run)
    den16 = cb.BatchDenoiser(16, model)
    o16, v16, taps = den16.process_streams(x[ids].contiguous(), unit_scale=True, return_taps=True)
    assert torch.equal(o16, out[ids]) and torch.equal(v16, vad[ids]), "a stream's result must not depend on its batch"
    taps = taps.cpu().numpy()
    flips = int((taps[:, :, 132].astype(np.int32) != rpi).sum())
    gain_diff = int(bits_differ(taps[:, :, 130], rpg).sum())
    sil_diff = int((taps[:, :, 133].astype(np.int32) != rsil).sum())
    # drift: the error of the last 10 s against that of the first 10 s
    e_first = float(np.abs(got[:, :1000 * 480] - ref[:, :1000 * 480]).max())
    e_last = float(np.abs(got[:, 5000 * 480:] - ref[:, 5000 * 480:]).max())
    decisions = {"pitch_index_flips": flips, "pitch_gain_bit_differences": gain_diff, "silence_gate_flips": sil_diff,
                 "max_abs_fs_first_10s": e_first, "max_abs_fs_last_10s": e_last, "silent_frame_fraction": float(rsil.mean()),
                 "frames_with_nan_pitch_gain_in_both": int((np.isnan(taps[:, :, 130]) & np.isnan(rpg)).sum())}
    assert flips == 0 and gain_diff == 0 and sil_diff == 0, decisions
    r = long_run_parity(ref, got, rvad, gv, rmargin, "c2")
    report("c2_16_of_1024_streams_x_6000_frames", {**decisions, **r})


def test_c4_ten_minute_meetings_i16_in_app_dual_mono(oracle_model, model):
    """configs[3]: 10-minute dual-source meetings, mic as PCM16, app audio f32, output dual-mono stereo PCM16, fed as
    ten 60 s calls.  (a) the mix against the oracle's denoiser + the reference's mixer/quantiser
    (commands/recording.rs:260-264, recording.rs:108-110) within 1e-3 FS + 1 LSB; (b) the quantiser itself bit-exact:
    the PCM16 the kernel writes equals trunc(clamp(f32 output + app) * 32767) of the kernel's own f32 output."""
    n, minutes = 4, 10
    calls, call_frames = minutes, 6000
    den_mix, den_f32 = cb.BatchDenoiser(n, model), cb.BatchDenoiser(n, model)
    mic_all, app_all, mix_all, f32_all, vad_all = [], [], [], [], []
    for c in range(calls):
        x = synth_device(n, call_frames, first_stream=40, start_frame=c * call_frames)
        mic = (x * 32767.0).round().clamp_(-32768, 32767).to(torch.int16)
        app = (torch.roll(x, 1, 0) * 0.5).contiguous()
        mix, _ = den_mix.process_streams(mic, unit_scale=True, app=app, mix_stereo_i16=True)
        o32, v32 = den_f32.process_streams(mic, unit_scale=True)
        vad_all.append(v32.cpu().numpy())
        mic_all.append(mic.cpu().numpy()), app_all.append(app.cpu().numpy())
        mix_all.append(mix.cpu().numpy()), f32_all.append(o32.cpu().numpy())
    mic, app = np.concatenate(mic_all, 1), np.concatenate(app_all, 1)
    mix, o32 = np.concatenate(mix_all, 1), np.concatenate(f32_all, 1)
    assert mix.shape == (n, minutes * 60 * 48000, 2) and np.array_equal(mix[:, :, 0], mix[:, :, 1])
    # (b) quantiser bit-exact on identical float input (-1 -> -32767 by truncation, recording.rs:498-501)
    mixed = np.clip(o32 + app, np.float32(-1.0), np.float32(1.0)).astype(np.float32)
    want_q = (mixed * np.float32(32767.0)).astype(np.int16)  # C-style truncation toward zero, as Rust `as i16`
    quant_diff = int((mix[:, :, 0] != want_q).sum())
    # (a) against the oracle
    ref, rvad, _, _, _, rmargin = po.process_streams_trace(oracle_model, mic.astype(np.float32) / np.float32(32768.0),
                                                           unit_scale=True, n_threads=n, native=True, margin=True)
    assert quant_diff == 0
    r = long_run_parity(ref, o32, rvad, np.concatenate(vad_all, 1), rmargin, "c4 denoised f32")
    # the PCM16 mix against the reference's mixer/quantiser fed with the oracle's output: 1e-3 FS = 32.8 LSB + 1 LSB of
    # quantisation, outside the frames long_run_parity sets apart (RNNoise's own discontinuity, tests/util.py)
    risky = rmargin < BRANCH_EPS
    risky[:, 1:] |= risky[:, :-1].copy()
    worst = 0
    for s in range(n):
        want = po.mix_dual_mono_i16(ref[s], app[s]).reshape(-1, 2)
        d = np.abs(mix[s].astype(np.int32) - want.astype(np.int32)).max(1).reshape(-1, 480).max(1)
        worst = max(worst, int(d[~risky[s]].max()))
    report("c4_4_meetings_x_10_min_i16_app_dual_mono", {
        "quantiser_mismatches_on_identical_float_input": quant_diff, "mix_max_abs_lsb_vs_oracle": worst, **r})
    assert worst <= 34


def test_c5_sixty_minute_streams_as_sixty_calls_with_state_carry(oracle_model, model):
    """configs[4]: 60-minute streams (360,000 frames) fed as sixty 60 s calls, every DenoiseState carried from call
    to call, against the oracle running each stream in one go."""
    n, calls, call_frames = CORES if CORES < 8 else 8, 60, 6000
    den = cb.BatchDenoiser(n, model)
    host_in = np.empty((n, calls * call_frames * 480), np.float32)
    host_out = np.empty_like(host_in)
    host_vad = np.empty((n, calls * call_frames), np.float32)
    host_pi = np.empty((n, calls * call_frames), np.int32)
    host_sil = np.empty((n, calls * call_frames), np.int32)
    for c in range(calls):
        x = synth_device(n, call_frames, first_stream=2, start_frame=c * call_frames)
        o, v, taps = den.process_streams(x, unit_scale=True, return_taps=True)
        sl = slice(c * call_frames * 480, (c + 1) * call_frames * 480)
        host_in[:, sl], host_out[:, sl] = x.cpu().numpy(), o.cpu().numpy()
        host_vad[:, c * call_frames:(c + 1) * call_frames] = v.cpu().numpy()
        host_pi[:, c * call_frames:(c + 1) * call_frames] = taps[:, :, 132].to(torch.int32).cpu().numpy()
        host_sil[:, c * call_frames:(c + 1) * call_frames] = taps[:, :, 133].to(torch.int32).cpu().numpy()
    assert den.frames_done == calls * call_frames
    ref, rvad, rpi, _, rsil, rmargin = po.process_streams_trace(oracle_model, host_in, unit_scale=True, n_threads=n,
                                                                native=True, margin=True)
    flips = int((host_pi != rpi).sum())
    sil_flips = int((host_sil != rsil).sum())
    # the worst frames outside the branch-margin set, for the record (stream, frame, error, margin, neighbourhood)
    ferr = np.abs(host_out.astype(np.float64) - ref).reshape(n, calls * call_frames, 480).max(2)
    risky = rmargin < BRANCH_EPS
    risky[:, 1:] |= risky[:, :-1].copy()
    masked = np.where(risky, 0.0, ferr)
    worst = []
    for idx in np.argsort(masked, axis=None)[::-1][:8]:
        s_, t_ = np.unravel_index(idx, masked.shape)
        lo = max(0, t_ - 3)
        worst.append({"stream": int(s_), "frame": int(t_), "err_fs": float(ferr[s_, t_]), "margin": float(rmargin[s_, t_]),
                      "vad": float(rvad[s_, t_]), "silence_around": rsil[s_, lo:t_ + 2].tolist(),
                      "err_around": [float(e) for e in ferr[s_, lo:t_ + 2]],
                      "margin_around": [float(e) for e in rmargin[s_, lo:t_ + 2]],
                      "pitch_around": rpi[s_, lo:t_ + 2].tolist()})
    report("c5_worst_frames_outside_the_branch_set", {"silence_gate_flips": sil_flips, "worst": worst})
    assert flips == 0 and sil_flips == 0
    r = long_run_parity(ref, host_out, rvad, host_vad, rmargin, "c5")
    last = slice(59 * call_frames * 480, None)
    report("c5_streams_x_60_min_as_60_calls", {
        "streams": n, "pitch_index_flips": flips, **r,
        "snr_db_last_minute": snr_db(ref[:, last], host_out[:, last]),
        "max_abs_fs_last_minute": float(np.abs(host_out[:, last] - ref[:, last]).max())})


def test_adversarial_inputs_on_the_gpu(oracle_model, model):
    """The adversarial set of tests/util.py through the CUDA kernels: pitch index, pitch gain and the silence gate
    equal the oracle's bit for bit; samples meet the tolerance except on the two inputs where RNNoise itself is
    discontinuous, where 2 % of full scale is allowed."""
    sigs = adversarial_signals(120)
    names = list(sigs)
    x = np.stack([sigs[k] for k in names])
    den = cb.BatchDenoiser(len(names), model)
    out, vad, taps = den.process_streams(torch.from_numpy(x).cuda(), unit_scale=False, return_taps=True)
    out, vad, taps = out.cpu().numpy(), vad.cpu().numpy(), taps.cpu().numpy()
    ref, rvad, rpi, rpg, rsil = po.process_streams_trace(oracle_model, x, n_threads=len(names))
    rep = {}
    for i, name in enumerate(names):
        r = parity_report(ref[i], out[i], rvad[i], vad[i])
        rep[name] = {"pitch_index_flips": int((taps[i, :, 132].astype(np.int32) != rpi[i]).sum()),
                     "pitch_gain_bit_differences": int(bits_differ(taps[i, :, 130], rpg[i]).sum()),
                     "silence_gate_flips": int((taps[i, :, 133].astype(np.int32) != rsil[i]).sum()),
                     "max_abs_fs": r["max_abs"] / 32768.0, "snr_db": r["snr_db"], "vad_max": r["vad_max"]}
    report("adversarial_120_frames", rep)
    for name, r in rep.items():
        assert r["pitch_index_flips"] == 0 and r["pitch_gain_bit_differences"] == 0 and r["silence_gate_flips"] == 0, (name, r)
        assert r["vad_max"] <= 1e-3, (name, r)
        if name in ("impulses", "square_clip"):
            assert r["max_abs_fs"] <= 0.02, (name, r)
        else:
            assert r["max_abs_fs"] <= 1e-3 and r["snr_db"] >= 60.0, (name, r)


def test_summation_order_flip_rate_report(oracle_model):
    """Not a GPU-vs-oracle check: the oracle against itself under another float32 summation order of the pitch
    path's inner products (rno_set_sum_policy), on 64 streams x 60 s generated on the device.  The decision-flip
    rate is reported separately from the sample error: a flip changes a whole frame."""
    x = synth_device(64, 6000, first_stream=128).cpu().numpy()
    base = po.process_streams_trace(oracle_model, x, unit_scale=True, n_threads=CORES, native=True, sum_policy=0)
    rep = {}
    for pol, what in ((1, "four partial sums in celt_inner_prod / dual_inner_prod"),
                      (2, "four partial sums there and in the cross-correlation kernels")):
        o, v, pi, pg, sil = po.process_streams_trace(oracle_model, x, unit_scale=True, n_threads=CORES, native=True, sum_policy=pol)
        fl = pi != base[2]
        rep[f"policy_{pol}"] = {"what": what, "frames": int(pi.size), "pitch_index_flips": int(fl.sum()),
                                "flip_rate": float(fl.mean()), "silence_gate_flips": int((sil != base[4]).sum()),
                                "frames_with_any_sample_over_1e-3_fs": int((np.abs(o - base[0]).reshape(64, 6000, 480).max(2) > 1e-3).sum()),
                                "max_abs_fs": float(np.abs(o - base[0]).max()), "snr_db": snr_db(base[0], o),
                                "vad_max": float(np.abs(v - base[1]).max())}
    report("oracle_summation_order_sensitivity_64x60s", rep)
    assert rep["policy_1"]["snr_db"] > 40 and rep["policy_2"]["snr_db"] > 40  # sanity only: this is a report


def test_c_abi_consumer_compiled_from_the_header(tmp_path):
    """tests/c_abi/abi_smoke.c: a C program compiled against include/crispy_ns.h and linked to libcrispy_ns.so runs
    create / process_frame / process_streams_host / multi / denoise_wav_files / destroy on the GPU."""
    exe = str(tmp_path / "abi_smoke")
    lib_dir = os.path.join(ROOT, "crispy_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi", "abi_smoke.c"), "-L" + lib_dir, "-lcrispy_ns", "-lm",
                           "-Wl,-rpath," + lib_dir, "-o", exe])
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    sys.stdout.write(r.stdout + r.stderr)
    assert r.returncode == 0 and "abi_smoke: ok" in r.stdout


def test_wav_files_in_dual_mono_files_out(oracle_model, model, tmp_path):
    """f3 end to end: eight stereo PCM16 recordings of different lengths (recording.rs:83-99 layout, written with the
    exact 44-byte header hound produces for 16-bit stereo) -> crispy_ns_denoise_wav_files -> dual-mono PCM16 files,
    against the oracle's denoiser + the reference's quantiser."""
    import struct
    n_files = 8
    x = synth_chunk(n_files, 48000 * 3, first_stream=300).numpy()
    paths_in, paths_out, lens = [], [], []
    for i in range(n_files):
        n = 48000 * 3 - 997 * i  # ragged: only file 0 is a whole number of frames
        q = (np.clip(x[i, :n], -1.0, 1.0) * np.float32(32767.0)).astype(np.int16)  # recording.rs:108-110
        data = np.repeat(q[:, None], 2, 1).tobytes()
        hdr = (b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 2, 48000, 48000 * 4, 4, 16)
               + b"data" + struct.pack("<I", len(data)))
        p = tmp_path / f"in{i}.wav"
        p.write_bytes(hdr + data)
        paths_in.append(str(p)), paths_out.append(str(tmp_path / f"out{i}.wav")), lens.append(n)
    mean_vad = cb.denoise_wav_files(paths_in, paths_out, model=model)
    worst = 0
    for i in range(n_files):
        y, sr = cb.wav_read_pcm16(paths_out[i])
        assert sr == 48000 and y.shape == (lens[i], 2) and np.array_equal(y[:, 0], y[:, 1])
        raw, _ = cb.wav_read_pcm16(paths_in[i])
        nfr = (lens[i] + 479) // 480
        xin = np.zeros(nfr * 480, np.float32)
        xin[:lens[i]] = raw[:, 0].astype(np.float32) / np.float32(32768.0)  # commands/transcription.rs:306-310
        ref, rv = po.process_streams(oracle_model, xin[None, :], unit_scale=True)
        want = po.mix_dual_mono_i16(ref[0], None).reshape(-1, 2)[:lens[i]]
        worst = max(worst, int(np.abs(y.astype(np.int32) - want.astype(np.int32)).max()))
        assert abs(mean_vad[i] - float(rv.mean())) <= 1e-3
    report("wav_files_8_ragged", {"max_abs_lsb_vs_oracle": worst})
    assert worst <= 34
    # first frame dropped (audio.rs:275-278): 480 samples shorter, same samples otherwise
    cb.denoise_wav_files(paths_in[:1], [str(tmp_path / "drop.wav")], model=model, drop_first_frame=True)
    yd, _ = cb.wav_read_pcm16(str(tmp_path / "drop.wav"))
    y0, _ = cb.wav_read_pcm16(paths_out[0])
    assert np.array_equal(yd, y0[480:])


def test_multi_device_handle_matches_single_batches(model):
    """crispy_ns_multi_*: contiguous stream blocks over the devices of the box (all of them; one here if the box has
    one), the same bits as one batch, PCM16 and f32."""
    ndev = cb.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0, 0]
    n = 11
    x = synth_chunk(n, 480 * 150, first_stream=7).pin_memory()
    md = cb.MultiDenoiser(n, devices, model)
    assert [r[1] for r in md.ranges] == [stream_block(n, len(devices), i)[0] for i in range(len(devices))]
    out, vad = md.process_streams_host(x, unit_scale=True)
    ref, rvad = cb.BatchDenoiser(n, model).process_streams_host(x, unit_scale=True)
    assert torch.equal(out, ref) and torch.equal(vad, rvad)
    md.reset()
    out2, _ = md.process_streams_host(x, unit_scale=True)
    assert torch.equal(out2, ref)
