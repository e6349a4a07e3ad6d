"""CPU tests of the oracle (oracle/rnnoise_oracle.c) against algorithm-level known answers and the
reference's own unit tests for the neighbouring rows.  PARITY UNPINNED vs nnnoiseless itself: the
crate is not in /root/reference and no Rust toolchain exists."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_half_window_power_complementary():
    # Vorbis window: w[i]^2 + w[479-i]^2 == 1 makes analysis+synthesis with 50% overlap an identity
    w = np.ctypeslib.as_array(po.lib().rno_half_window(), shape=(480,)).astype(np.float64)
    assert np.allclose(w ** 2 + w[::-1] ** 2, 1.0, atol=2e-7)
    assert np.all(np.diff(w) >= 0) and w[0] > 0 and w[-1] <= 1.0


def test_dct_table_orthonormal():
    d = np.ctypeslib.as_array(po.lib().rno_dct_table(), shape=(22, 22)).astype(np.float64)
    m = d * np.sqrt(2.0 / 22)
    assert np.allclose(m.T @ m, np.eye(22), atol=1e-6)


def test_tansig_table_and_approx():
    t = np.ctypeslib.as_array(po.lib().rno_tansig_table(), shape=(201,))
    ref = np.tanh(0.04 * np.arange(201))
    assert np.max(np.abs(t - ref)) <= 5.7e-7  # 6 printed decimals, then f32
    assert t[0] == 0.0 and abs(t[25] - 0.761594) < 1e-6
    xs = np.linspace(-9, 9, 2001).astype(np.float32)
    y = np.array([po.lib().rno_tansig_approx(float(x)) for x in xs])
    assert np.max(np.abs(y - np.tanh(xs))) < 3e-4
    s = np.array([po.lib().rno_sigmoid_approx(float(x)) for x in xs])
    assert np.max(np.abs(s - 1 / (1 + np.exp(-xs.astype(np.float64))))) < 2e-4
    assert po.lib().rno_tansig_approx(8.0) == 1.0 and po.lib().rno_tansig_approx(-8.0) == -1.0


def test_transforms_match_numpy():
    rng = np.random.default_rng(1)
    x = rng.standard_normal(960).astype(np.float32) * 1000
    X = po.forward_transform(x)
    ref = np.fft.rfft(x.astype(np.float64)) / 960  # forward carries 1/960
    assert np.max(np.abs(X - ref)) < 2e-6 * np.max(np.abs(ref))
    xi = po.inverse_transform(X)  # inverse carries no scaling -> identity overall
    assert np.max(np.abs(xi - x)) < 1e-3


def _biquad_f64(x):
    a0, a1 = float(np.float32(-1.99599)), float(np.float32(0.99600))
    m0 = m1 = 0.0
    y = np.empty_like(x, dtype=np.float64)
    for i, xi in enumerate(x.astype(np.float64)):
        yi = xi + m0
        m0 = m1 + (-2.0 * xi - a0 * yi)
        m1 = xi - a1 * yi
        y[i] = yi
    return y


def test_silence_path_is_delayed_highpass(oracle_model):
    # E < 0.04: no gains, no pitch filter -> out = biquad(in) delayed by one frame (perfect
    # reconstruction of the windowed overlap-add), VAD = 0, RNN state untouched.
    rng = np.random.default_rng(2)
    x = (rng.standard_normal(480 * 12) * 0.02).astype(np.float32)
    out, vad = po.process_streams(oracle_model, x[None, :])
    assert np.all(vad == 0.0)
    hp = _biquad_f64(x)
    # upstream keeps the biquad memory in f32, which drifts ~1e-4 relative against an f64 recursion
    assert np.max(np.abs(out[0, 480:] - hp[:-480])) < 2e-5
    assert np.max(np.abs(out[0, :480])) < 1e-6


def test_model_blob_roundtrip_and_text_format(oracle_model):
    blob = oracle_model.to_bytes()
    assert blob[:8] == b"CRNSMDL1" and len(blob) == 8 + 6 * 16 + 87503
    again = po.Model.from_bytes(blob)
    assert again.to_bytes() == blob
    with pytest.raises(ValueError):
        po.Model.from_bytes(blob[:-1])
    # rnnoise-nu text format: header, then per layer "in out act" + integers; layer order
    # input_dense, vad_gru, noise_gru, denoise_gru, denoise_output, vad_output
    def layers(b):
        off, out = 8, []
        for _ in range(6):
            kind, i, n, act = np.frombuffer(b[off:off + 16], dtype="<u4")
            off += 16
            cnt = i * n + n if kind == 0 else i * 3 * n + n * 3 * n + 3 * n
            out.append((int(i), int(n), int(act), np.frombuffer(b[off:off + cnt], dtype=np.int8)))
            off += int(cnt)
        return out
    L = layers(blob)
    order = [0, 1, 3, 4, 5, 2]
    text = "rnnoise-nu model file version 1\n"
    for k in order:
        i, n, act, w = L[k]
        text += f"{i} {n} {act}\n" + " ".join(str(int(v)) for v in w) + "\n"
    assert po.Model.from_bytes(text.encode()).to_bytes() == blob


def test_chunk_invariance_and_determinism(oracle_model):
    from tests.util import make_signal
    x = make_signal(1, 40)[0]
    full, _ = po.process_streams(oracle_model, x[None, :])
    st = po.DenoiseState(oracle_model)
    pieces = [st.process_frame(x[t * 480:(t + 1) * 480])[0] for t in range(40)]
    assert np.array_equal(np.concatenate(pieces), full[0])


# ---- the reference's own unit tests for the neighbouring rows ------------------------------------
def test_linear_resampler_same_rate_passthrough():  # audio.rs:1041-1053
    x = (np.arange(10) * 0.1).astype(np.float32)
    y = po.linear_resample(x, 48000.0, 48000.0)
    assert len(y) == 10 and np.max(np.abs(y - x)) < 0.001


def test_linear_resampler_downsample_produces_fewer():  # audio.rs:1055-1067
    y = po.linear_resample(np.full(300, 0.5, np.float32), 48000.0, 16000.0)
    assert 80 < len(y) < 120


def test_linear_resampler_upsample_produces_more():  # audio.rs:1069-1081
    y = po.linear_resample(np.full(100, 0.5, np.float32), 16000.0, 48000.0)
    assert 250 < len(y) < 350
    assert np.all(y == 0.5)


def test_linear_resampler_441_to_48_is_linear_interp():
    n = 4410
    x = np.sin(2 * np.pi * 440 * np.arange(n) / 44100).astype(np.float32)
    y = po.linear_resample(x, 44100.0, 48000.0)
    step = float(np.float32(44100.0) / np.float32(48000.0))
    pos = np.arange(len(y)) * step
    ref = np.interp(pos, np.arange(n), x.astype(np.float64))
    assert abs(len(y) - round((n - 1) / step)) <= 1
    assert np.max(np.abs(y - ref)) < 1e-5


def test_mixer_quantiser_matches_wavwriter_tests():  # recording.rs:454-504
    q = po.mix_dual_mono_i16(np.array([0.5, -0.5], np.float32), None)
    assert q.tolist() == [16383, 16383, -16383, -16383]  # (0.5*32767) as i16 truncates
    q = po.mix_dual_mono_i16(np.array([2.0, -3.0, 1.5, -1.5], np.float32), None)
    assert q.tolist() == [32767, 32767, -32767, -32767, 32767, 32767, -32767, -32767]
    q = po.mix_dual_mono_i16(np.array([0.25], np.float32), np.array([0.5], np.float32))
    assert q.tolist() == [int(0.75 * 32767)] * 2  # commands/recording.rs:260-264: mixed to both channels


def test_processor_wrapper_matches_manual_composition(oracle_model):  # audio.rs:242-295
    from tests.util import make_signal
    x = make_signal(1, 12)[0] / 32768.0
    y = po.processor_run(oracle_model, x.astype(np.float32), 48000.0, 0.8)
    assert len(y) == 11 * 480  # first frame dropped (audio.rs:275-278)
    ref, _ = po.process_streams(oracle_model, x[None, :].astype(np.float32), unit_scale=True, volume=0.8)
    assert np.array_equal(y, ref[0, 480:])
    # 44.1 kHz input goes through LinearResampler first (audio.rs:217-221)
    x44 = x[: 441 * 10].astype(np.float32)
    y44 = po.processor_run(oracle_model, x44, 44100.0, 1.0)
    rs = po.linear_resample(x44, 44100.0, 48000.0)
    nfr = len(rs) // 480
    ref, _ = po.process_streams(oracle_model, rs[None, : nfr * 480], unit_scale=True)
    assert np.array_equal(y44, ref[0, 480:])


def test_golden_regression(oracle_model):
    """The committed fixture was produced by this oracle (tests/golden/make_golden.py); it pins the
    oracle against accidental change -- it is NOT an nnnoiseless output (none can be made here)."""
    g = np.load(os.path.join(GOLDEN, "c1_head.npz"))
    x = g["x_i16"].astype(np.float32)
    out, vad = po.process_streams(oracle_model, x[None, :])
    assert np.array_equal(out[0], g["out"])
    assert np.array_equal(vad[0], g["vad"])
    _, taps = po.debug_trace(oracle_model, x)
    assert [t["pitch_index"] for t in taps] == g["pitch_index"].tolist()


# ---- f2 / north_star item 4: windowed-sinc front end (rubato-equivalent; rubato 0.16.2, Cargo.lock:4166) ----
def test_sinc_table_is_a_normalised_symmetric_lowpass():
    h, M = po.sinc_table(44100, 48000)
    assert h.shape == (160, 256) and M == 147
    # make_sincs normalisation: all taps sum to the oversampling factor, every phase has ~unit DC gain
    assert abs(float(h.astype(np.float64).sum()) - 160.0) < 1e-3
    assert np.abs(h.astype(np.float64).sum(axis=1) - 1.0).max() < 1e-4
    # phase 0 is symmetric about the interpolation point; phase p mirrors phase L-p
    assert np.allclose(h[0, :255], h[0, :255][::-1], atol=1e-9)
    assert np.allclose(h[1, :], h[159, ::-1], atol=1e-9)
    # cutoff 0.95 x Nyquist of the input: the prototype passes 0.8 and stops 1.15 (units of input Nyquist)
    proto = np.zeros(160 * 256)
    for p in range(160):
        proto[160 * (np.arange(256) + 1) - p - 1] = h[p]
    w = np.fft.rfftfreq(1 << 18) * 2 * 160  # in units of the input Nyquist
    H = np.abs(np.fft.rfft(proto, 1 << 18)) / 160.0
    assert np.abs(H[w < 0.8] - 1.0).max() < 1e-3
    assert H[(w > 1.15) & (w < 20)].max() < 1e-5


def test_sinc_resample_counts_and_frame_alignment():
    # 441 samples at 44.1 kHz are exactly one 480-sample frame at 48 kHz
    for frames in (1, 2, 7):
        assert len(po.sinc_resample(np.zeros(441 * frames, np.float32), 44100, 48000)) == 480 * frames
    assert len(po.sinc_resample(np.zeros(300, np.float32), 48000, 16000)) == 100
    assert len(po.sinc_resample(np.zeros(100, np.float32), 16000, 48000)) == 300
    assert len(po.sinc_resample(np.zeros(0, np.float32), 44100, 48000)) == 0
    assert len(po.sinc_resample(np.zeros(1, np.float32), 44100, 48000)) == 2


def test_sinc_resample_matches_scipy_polyphase():
    from scipy.signal import upfirdn
    rng = np.random.default_rng(5)
    x = rng.standard_normal(3000).astype(np.float32)
    for rin, rout in ((44100, 48000), (48000, 44100), (16000, 48000), (48000, 16000)):
        h, M = po.sinc_table(rin, rout)
        L = h.shape[0]
        proto = np.zeros(L * 256)
        for p in range(L):
            proto[L * (np.arange(256) + 1) - p - 1] = h[p]  # tap at x = L*(k+1) - p, stored from x = 1
        # out[n] = sum_i x[i] * g(n*M - i*L) with g centred at x = L*128: upfirdn delay = L*128 - 1
        full = upfirdn(proto, x.astype(np.float64), up=L, down=1)
        y = po.sinc_resample(x, rin, rout)
        ref = full[L * 128 - 1 + M * np.arange(len(y))]
        assert np.abs(y - ref).max() < 2e-5 * max(1.0, np.abs(ref).max()), (rin, rout)


def test_sinc_resample_reconstructs_a_tone_and_beats_linear():
    n = 44100
    t = np.arange(n) / 44100.0
    x = (0.5 * np.sin(2 * np.pi * 5000.0 * t)).astype(np.float32)
    y = po.sinc_resample(x, 44100, 48000)
    want = 0.5 * np.sin(2 * np.pi * 5000.0 * np.arange(len(y)) / 48000.0)
    mid = slice(400, len(y) - 400)
    err_sinc = np.abs(y[mid] - want[mid]).max()
    assert err_sinc < 2e-5  # f32 accumulation over 256 taps
    lin = po.linear_resample(x, 44100.0, 48000.0)
    k = min(len(lin), len(want)) - 400
    # the reference's linear interpolator (audio.rs:108-133) starts one input sample late and smears a 5 kHz tone
    best = min(np.abs(lin[400:k] - 0.5 * np.sin(2 * np.pi * 5000.0 * (np.arange(400, k) / 48000.0 + d / 44100.0))).max()
               for d in (0.0, 1.0))
    assert best > 100 * err_sinc


def test_front_end_golden_regression():
    """tests/golden/front_end.npz (make_golden.py): the linear and sinc resamplers of the oracle on a chirp."""
    g = np.load(os.path.join(GOLDEN, "front_end.npz"))
    assert np.array_equal(po.linear_resample(g["x44"], 44100.0, 48000.0), g["linear"])
    assert np.array_equal(po.sinc_resample(g["x44"], 44100, 48000), g["sinc"])
    h, _ = po.sinc_table(44100, 48000)
    assert np.array_equal(h[0], g["sinc_taps_phase0"]) and np.array_equal(h[77], g["sinc_taps_phase77"])



def test_pitch_path_matches_an_independent_numpy_transliteration(oracle_model):
    """tests/np_pitch.py restates biquad + pitch_downsample + pitch_search + remove_doubling from the published
    algorithm (xiph/rnnoise pitch.c / celt_lpc.c), independently of the C oracle: the discrete
    pitch decisions and the pitch gain of every frame must agree bit for bit."""
    from tests.np_pitch import PitchTracker
    from tests.util import make_signal
    x = make_signal(4, 150)
    for s in range(4):
        _, taps = po.debug_trace(oracle_model, x[s].astype(np.float32))
        pt = PitchTracker()
        for t in range(150):
            pi, g = pt.frame(x[s, t * 480:(t + 1) * 480])
            assert pi == taps[t]["pitch_index"], (s, t)
            assert np.float32(g) == np.float32(taps[t]["pitch_gain"]), (s, t, float(g), float(taps[t]["pitch_gain"]))


def test_whole_frame_matches_an_independent_numpy_transliteration(oracle_model):
    """tests/np_denoise.py restates the rest of process_frame (analysis, features, dense/GRU stack, pitch filter,
    synthesis) in float64 from the published algorithm; the float32 C oracle must follow it: same silence decisions,
    band energies / features / gains / VAD to float32 accuracy, output within 3e-5 of full scale and >= 80 dB (float32
    against float64 through 150 recurrent steps)."""
    from tests.np_denoise import NpDenoise
    from tests.util import make_signal, snr_db
    x = make_signal(2, 150)  # stream 1 carries a digital-silence stretch
    blob = oracle_model.to_bytes()
    for s in range(2):
        out_o, taps = po.debug_trace(oracle_model, x[s].astype(np.float32))
        ref = NpDenoise(blob)
        outs, n_silent = [], 0
        for t in range(150):
            o, vad, tp = ref.process_frame(x[s, t * 480:(t + 1) * 480])
            ot = taps[t]
            outs.append(o)
            assert tp["pitch_index"] == ot["pitch_index"] and tp["silence"] == ot["silence"], (s, t)
            n_silent += tp["silence"]
            assert np.allclose(tp["Ex"], ot["Ex"], rtol=1e-4, atol=1e-6)
            assert np.allclose(tp["Ep"], ot["Ep"], rtol=1e-4, atol=1e-6)
            assert np.abs(tp["Exp"] - np.array(ot["Exp"])).max() < 1e-3
            assert np.abs(tp["features"] - np.array(ot["features"])).max() < 1e-3
            assert abs(vad - float(ot["vad"])) < 1e-5
            if not tp["silence"]:
                assert np.abs(tp["gains"] - np.array(ot["gains"])).max() < 1e-4
        out = np.concatenate(outs)
        assert np.abs(out - out_o).max() < 1.0 and snr_db(out, out_o) >= 80.0
        assert (n_silent > 0) == (s == 1)



def test_pitch_path_numpy_cross_check_on_adversarial_inputs(oracle_model):
    """The same bit-for-bit agreement on inputs that reach the corners of the pitch range (60 .. 767)."""
    from tests.np_pitch import PitchTracker
    from tests.util import adversarial_signals
    seen = set()
    for name, x in adversarial_signals(40).items():
        _, taps = po.debug_trace(oracle_model, x)
        pt = PitchTracker()
        for t in range(40):
            pi, g = pt.frame(x[t * 480:(t + 1) * 480])
            assert pi == taps[t]["pitch_index"] and np.float32(g) == np.float32(taps[t]["pitch_gain"]), (name, t)
            seen.add(pi)
    assert min(seen) == 60 and max(seen) > 700


def test_summation_policy_switch_and_decision_trace(oracle_model):
    """rno_set_sum_policy: 0 (default) is the sequential xiph order; 1 and 2 regroup the inner products into four
    partial sums.  The trace variant returns the same samples as the plain call plus the per-frame decisions; another
    summation order keeps the output within tolerance on ordinary frames and flips only a small share of the pitch
    decisions (the rate at full size is in profiles/r2_parity.json)."""
    from tests.util import make_signal, snr_db
    x = make_signal(4, 400)
    o0, v0 = po.process_streams(oracle_model, x, n_threads=4)
    o, v, pi, pg, sil = po.process_streams_trace(oracle_model, x, n_threads=4)
    assert np.array_equal(o, o0) and np.array_equal(v, v0)
    _, taps = po.debug_trace(oracle_model, x[1])
    assert pi[1].tolist() == [t["pitch_index"] for t in taps] and sil[1].tolist() == [t["silence"] for t in taps]
    assert np.array_equal(pg[1], np.array([t["pitch_gain"] for t in taps], np.float32))
    assert po.lib().rno_get_sum_policy() == 0
    for pol in (1, 2):
        o2, v2, pi2, pg2, sil2 = po.process_streams_trace(oracle_model, x, n_threads=4, sum_policy=pol)
        assert po.lib().rno_get_sum_policy() == 0  # restored
        assert not np.array_equal(pg2, pg)  # another rounding order gives other correlations (last bits of the gain) ...
        assert (pi2 != pi).mean() < 0.01 and np.array_equal(sil2, sil)  # ... but nearly every decision stands; the
        # samples only change where a decision flips (the inner products feed nothing but decisions)
        assert snr_db(o, o2) > 50
    # the native (-O3 -march=native) build used by the CPU baseline computes the same bits as the checker build
    po.build_native()
    on, vn = po.process_streams(oracle_model, x, n_threads=4, native=True)
    assert np.array_equal(on, o0) and np.array_equal(vn, v0)


def test_branch_margin_names_the_frames_the_pitch_filter_makes_irreproducible(oracle_model):
    """RNNoise's pitch filter is discontinuous at Exp == g (r jumps to 1).  With the pitch filter's inputs perturbed
    the way another float32 implementation perturbs them (Exp by 1e-6 and a relative 1e-5, the band gains by 1e-4 in the
    logit domain; nothing that feeds the state), the output moves by ~1 % of full scale on a few frames and by < 1e-4 on all
    others -- and the frames that move are inside the set the oracle's branch margin flags (with their successors:
    overlap-add), which stays a fraction of a per cent.  This is the criterion the long-run GPU parity tests use
    (tests/util.py long_run_parity); tests/diag/pitch_filter_conditioning.py is the long version."""
    import torch
    from crispy_b200.synth import synth_chunk
    from tests.util import BRANCH_EPS
    n, nf = 4, 12000
    x = torch.cat([synth_chunk(n, 6000 * 480, first_stream=2, start_sample=c * 6000 * 480) for c in range(nf // 6000)], 1).numpy()
    ref, _, _, _, _, mg = po.process_streams_trace(oracle_model, x, unit_scale=True, n_threads=4, native=True, margin=True)
    risky = mg < BRANCH_EPS
    risky[:, 1:] |= risky[:, :-1].copy()
    assert 0 < risky.mean() < 0.01
    L = po.lib(True)
    moved = 0
    for d_exp, d_g, a_exp in ((1e-5, -1e-4, 1e-6), (-1e-5, 1e-4, -1e-6)):
        L.rno_set_pf_perturb(d_exp, d_g, a_exp)
        try:
            out2 = po.process_streams_trace(oracle_model, x, unit_scale=True, n_threads=4, native=True)[0]
        finally:
            L.rno_set_pf_perturb(0.0, 0.0, 0.0)
        err = np.abs(out2.astype(np.float64) - ref).reshape(n, nf, 480).max(2)
        assert err[~risky].max() <= 3e-4, (d_exp, d_g, a_exp, err[~risky].max())
        moved += int((err > 3e-4).sum())
    again = po.process_streams_trace(oracle_model, x, unit_scale=True, n_threads=4, native=True)[0]
    assert np.array_equal(again, ref)  # the hook is off again
    print("frames moved by > 3e-4 FS under the perturbations:", moved, "flagged:", int(risky.sum()), "of", risky.size)


def _np_resample_audio(x, from_rate, to_rate):
    """recording.rs:13-39 transliterated with NumPy (f64 positions, f32 samples), independent of the C oracle."""
    x = np.asarray(x, np.float32)
    if from_rate == to_rate:
        return x.copy()
    ratio = np.float64(from_rate) / np.float64(to_rate)
    n_out = int(np.ceil(np.float64(len(x)) / ratio))
    pos = np.arange(n_out, dtype=np.float64) * ratio
    idx = np.floor(pos).astype(np.int64)
    frac = (pos - idx).astype(np.float32)
    keep = idx < len(x)
    idx, frac = idx[keep], frac[keep]
    nxt = np.minimum(idx + 1, len(x) - 1)
    two = idx + 1 < len(x)
    out = x[idx] + (x[nxt] - x[idx]) * frac
    return np.where(two, out, x[idx]).astype(np.float32)


def test_resample_audio_follows_the_recorder():  # recording.rs:13-39 (app audio ahead of the dual-mono mix)
    rng = np.random.default_rng(7)
    for n, fr, to in ((0, 44100, 48000), (1, 44100, 48000), (2, 44100, 48000), (441, 44100, 48000), (44101, 44100, 48000),
                      (4800, 48000, 44100), (1000, 16000, 48000), (999, 48000, 16000), (777, 48000, 48000), (12345, 22050, 48000)):
        x = rng.standard_normal(n).astype(np.float32)
        got, want = po.resample_audio(x, fr, to), _np_resample_audio(x, fr, to)
        assert got.shape == want.shape and np.array_equal(got, want), (n, fr, to)
    # known answers: identical rates copy; a ramp stays a ramp of slope from/to; 441 samples make one 480-sample frame
    x = np.arange(100, dtype=np.float32)
    assert np.array_equal(po.resample_audio(x, 48000, 48000), x)
    y = po.resample_audio(x, 44100, 48000)
    assert len(y) == int(np.ceil(100 / (44100 / 48000))) and np.allclose(y[:-1], np.arange(len(y) - 1) * (44100 / 48000), atol=1e-4)
    assert y[-1] == x[-1]  # the last output sits on the last sample alone (recording.rs:32-35)
    assert len(po.resample_audio(np.zeros(441, np.float32), 44100, 48000)) == 480


def test_downmix_mono_follows_the_capture_callbacks():  # audio.rs:754-755 (f32), :816-818 (i16), :879-884 (u16)
    rng = np.random.default_rng(3)
    for ch in (1, 2, 3, 6):
        xf = rng.standard_normal(50 * ch).astype(np.float32)
        want = np.zeros(50, np.float32)
        for c in range(ch):  # the f32 sum in channel order, from zero
            want = (want + xf[c::ch]).astype(np.float32)
        assert np.array_equal(po.downmix_mono(xf, ch), (want / np.float32(ch)).astype(np.float32))
        xi = rng.integers(-32768, 32768, 50 * ch).astype(np.int16)
        want = np.zeros(50, np.float32)
        for c in range(ch):
            want = (want + xi[c::ch].astype(np.float32) / np.float32(32768.0)).astype(np.float32)
        assert np.array_equal(po.downmix_mono(xi, ch), (want / np.float32(ch)).astype(np.float32))
        xu = rng.integers(0, 65536, 50 * ch).astype(np.uint16)
        want = np.zeros(50, np.float32)
        for c in range(ch):
            want = (want + (xu[c::ch].astype(np.float32) - np.float32(32768.0)) / np.float32(32768.0)).astype(np.float32)
        assert np.array_equal(po.downmix_mono(xu, ch), (want / np.float32(ch)).astype(np.float32))
    assert po.downmix_mono(np.array([32767, 32767], np.int16), 2)[0] == np.float32(32767 / 32768)
    assert po.downmix_mono(np.array([0, 65535], np.uint16), 2)[0] == np.float32(-1 / 65536)
