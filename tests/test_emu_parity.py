"""CPU: the pipeline kernels' own code (crispy_b200/csrc/ns_pipe.cuh) run under the host SIMT
emulation (tests/emu) against the oracle.  Sizes are small because every CUDA thread is an OS thread."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import (TOL_MAX_ABS, assert_parity, canonical_state, emu_process, snr_db, make_signal)

F_UNIT, F_OUT_I16, F_IN_I16, F_MIX = 4, 2, 1, 8


@pytest.fixture(scope="module")
def sig():
    return make_signal(2, 20)


def test_emulated_kernel_matches_oracle(oracle_model, model_blob, sig):
    out, vad, dbg, _ = emu_process(model_blob, sig, chunk=8)
    ref, rvad = po.process_streams(oracle_model, sig)
    r = assert_parity(ref, out, rvad, vad, "emu chunk=8")
    # the recurrent core feeds the tensor pipe bf16 hi+lo activations (2^-17 relative): ~2e-6 FS
    assert r["snr_db"] > 95.0 and r["max_abs"] < 0.5
    # stage taps of stream 1 (contains a digital-silence stretch)
    _, taps = po.debug_trace(oracle_model, sig[1])
    sil = np.array([t["silence"] for t in taps])
    assert sil[:4].sum() == 4 and sil[4:].sum() == 0, "leading digital silence must trip the E < 0.04 gate"
    assert np.array_equal(dbg[1, :, 133].astype(int), sil)
    # the pitch decision chain (biquad, pitch_downsample, pitch_search, remove_doubling) is bit-exact
    assert np.array_equal(dbg[1, :, 132].astype(int), np.array([t["pitch_index"] for t in taps]))
    assert np.array_equal(dbg[1, :, 130], np.array([t["pitch_gain"] for t in taps], dtype=np.float32))
    feats = np.array([t["features"] for t in taps])
    assert np.max(np.abs(dbg[1, :, 0:42] - feats)) < 1e-4
    gains = np.array([t["gains"] for t in taps])
    assert np.max(np.abs(dbg[1, :, 42:64] - gains)) < 1e-4
    ex = np.array([t["Ex"] for t in taps])
    rel = np.abs(dbg[1, :, 64:86] - ex) / (ex.max(axis=1, keepdims=True) + 1.0)
    assert rel.max() < 1e-5


def test_chunk_size_does_not_change_results(model_blob, sig):
    a, va, _, sa = emu_process(model_blob, sig, chunk=20)  # one chunk
    b, vb, _, sb = emu_process(model_blob, sig, chunk=3)   # seven chunks, runs shorter than the pitch run
    c, vc, _, sc = emu_process(model_blob, sig, chunk=1)   # frame by frame (history shorter than kHist per chunk)
    sa, sb, sc = canonical_state(sa), canonical_state(sb), canonical_state(sc)
    assert np.array_equal(a, b) and np.array_equal(va, vb) and np.array_equal(sa, sb)
    assert np.array_equal(a, c) and np.array_equal(va, vc) and np.array_equal(sa, sc)


def test_chunked_state_carry_is_bit_exact(model_blob, sig):
    full, vfull, _, st_full = emu_process(model_blob, sig, chunk=8)
    o1, v1, _, st = emu_process(model_blob, sig[:, : 7 * 480], chunk=8)
    o2, v2, _, st = emu_process(model_blob, sig[:, 7 * 480:], chunk=8, state=st)
    assert np.array_equal(np.concatenate([o1, o2], axis=1), full)
    assert np.array_equal(np.concatenate([v1, v2], axis=1), vfull)
    assert np.array_equal(canonical_state(st), canonical_state(st_full))


def test_ragged_stream_count_and_single_frame(oracle_model, model_blob):
    x = make_signal(3, 3)  # 3 streams: the biquad warp and the 8-stream RNN CTA are mostly empty
    out, vad, _, _ = emu_process(model_blob, x, chunk=2)
    ref, rvad = po.process_streams(oracle_model, x)
    assert_parity(ref, out, rvad, vad, "ragged")
    one, v1, _, _ = emu_process(model_blob, x[:1, :480], chunk=1)
    assert_parity(ref[:1, :480], one, rvad[:1, :1], v1, "single frame")


def test_wrapper_flags_unit_scale_i16_and_mix(oracle_model, model_blob, sig):
    xu = (sig / 32768.0).astype(np.float32)
    ref, rvad = po.process_streams(oracle_model, xu, unit_scale=True, volume=0.7)
    out, vad, _, _ = emu_process(model_blob, xu, chunk=8, flags=F_UNIT, volume=0.7)
    assert np.max(np.abs(out - ref)) <= TOL_MAX_ABS / 32768.0 and snr_db(ref, out) > 60
    # int16 in / int16 out (16-bit scale, round to nearest)
    xi = np.clip(np.rint(sig), -32768, 32767).astype(np.int16)
    ref16, _ = po.process_streams(oracle_model, xi.astype(np.float32))
    o16, _, _, _ = emu_process(model_blob, xi, chunk=8, flags=F_IN_I16 | F_OUT_I16, out_dtype=np.int16)
    assert np.max(np.abs(o16.astype(np.float64) - np.rint(ref16))) <= TOL_MAX_ABS + 1
    # f1: dual-mono mix with app audio, PCM16 truncation (recording.rs:108-110)
    rng = np.random.default_rng(3)
    app = (rng.standard_normal(xu.shape) * 0.1).astype(np.float32)
    mix, _, _, _ = emu_process(model_blob, xu, chunk=8, flags=F_UNIT | F_MIX, volume=1.0, out_dtype=np.int16,
                               out_cols=2 * xu.shape[1], app=app)
    refd, _ = po.process_streams(oracle_model, xu, unit_scale=True, volume=1.0)
    for s in range(2):
        want = po.mix_dual_mono_i16(refd[s], app[s])
        assert np.max(np.abs(mix[s].astype(np.int32) - want.astype(np.int32))) <= 34  # 1e-3 FS + 1 LSB
        assert np.array_equal(mix[s][0::2], mix[s][1::2])


def test_drop_first_frame_offset(model_blob, sig):
    full, _, _, _ = emu_process(model_blob, sig, chunk=8)
    dropped, _, _, _ = emu_process(model_blob, sig, chunk=8, out_frame_offset=-1)
    assert np.array_equal(dropped[:, : 19 * 480], full[:, 480:])


def test_adversarial_inputs_keep_every_decision_bit_exact(oracle_model, model_blob):
    """DC, impulses, loud / tiny noise, tones at and beyond the pitch range, clipping, a chirp, silence then a step:
    pitch index, pitch gain and the silence gate equal the oracle's bit for bit, features / gains / VAD to float32
    accuracy.  Samples meet the north_star tolerance except on two inputs where RNNoise itself is discontinuous:
    a clipped square wave drives the band correlation Exp and the gain g both to 1.0, where the pitch filter's
    `Exp > g ? 1 : ...` branch flips on the last bit, and sparse impulses leave bands of the lagged window at
    rounding-noise energy under `Ex / (1e-8 + Ep)`.  There two correct implementations differ by ~1 % of full scale
    in isolated frames -- the float32 oracle and the float64 NumPy transliteration do, too."""
    from tests.util import adversarial_signals, parity_report
    sigs = adversarial_signals(24)
    names = list(sigs)
    x = np.stack([sigs[k] for k in names])
    out, vad, dbg, _ = emu_process(model_blob, x, chunk=8)
    ref, rvad = po.process_streams(oracle_model, x)
    for i, name in enumerate(names):
        _, taps = po.debug_trace(oracle_model, x[i])
        assert np.array_equal(dbg[i, :, 132].astype(int), np.array([t["pitch_index"] for t in taps])), name
        assert np.array_equal(dbg[i, :, 130], np.array([t["pitch_gain"] for t in taps], dtype=np.float32)), name
        assert np.array_equal(dbg[i, :, 133].astype(int), np.array([t["silence"] for t in taps])), name
        assert np.max(np.abs(dbg[i, :, 0:42] - np.array([t["features"] for t in taps]))) < 2e-2, name  # log-domain
        assert np.max(np.abs(dbg[i, :, 42:64] - np.array([t["gains"] for t in taps]))) < 5e-3, name
        r = parity_report(ref[i], out[i], rvad[i], vad[i])
        assert r["vad_max"] <= 1e-3, name
        if name in ("impulses", "square_clip"):
            assert r["max_abs"] <= 0.02 * 32768.0, (name, r)
        else:
            assert r["max_abs"] <= TOL_MAX_ABS and r["snr_db"] >= 60.0, (name, r)


def test_second_generation_pitch_kernel_is_bit_exact_too(oracle_model, model_blob):
    """ns_pitch7.cuh (-DNS_PITCH_V7: coarse search approximated on the tensor pipe, bracketed, and recomputed exactly
    only for the lags that can win) names the same pitch index and gain as the oracle on speech, on silence after
    speech (the exact path: the signal decays through the denormal range) and on the adversarial set."""
    from tests.util import adversarial_signals
    x = make_signal(2, 96)
    x[:, 60 * 480:] = 0.0
    adv = adversarial_signals(16)
    for name, sig in (("speech then zeros", x), ("adversarial", np.stack(list(adv.values())))):
        out, vad, dbg, _ = emu_process(model_blob, sig, chunk=8, v7=True)
        ref, rvad, rpi, rpg, rsil = po.process_streams_trace(oracle_model, sig, n_threads=4)
        assert np.array_equal(dbg[:, :, 132].astype(np.int32), rpi), name
        assert not ((dbg[:, :, 130] != rpg) & ~(np.isnan(dbg[:, :, 130]) & np.isnan(rpg))).any(), name
        assert np.array_equal(dbg[:, :, 133].astype(np.int32), rsil), name


def _biquad_reference(x, m0, m1):
    """upstream denoise.c biquad(): f32 state, f64 intermediates (oracle/rnnoise_oracle.c biquad, line for line)"""
    a0, a1 = float(np.float32(-1.99599)), float(np.float32(0.99600))
    y = np.empty_like(x)
    m0, m1 = np.float32(m0), np.float32(m1)
    for i, xi in enumerate(x):
        yi = np.float32(xi + m0)
        m0n = np.float32(float(m1) + (-2.0 * float(xi) - a0 * float(yi)))
        m1 = np.float32(float(xi) - a1 * float(yi))
        m0 = m0n
        y[i] = yi
    return y, m0, m1


def test_speculative_biquad_recomputes_what_it_misses(model_blob):
    """K0 runs the recursion speculatively in error-free f32 arithmetic on one warp, and upstream's f64 expression on
    four others, one 24-sample segment each, from the states the first recorded; a segment that ends off the record
    makes the first warp repair the tile and speculate the next one again.  A filter state decaying into digital
    silence crosses 2^-126, where the low half of a0*y underflows and the speculation goes wrong a few hundred times;
    a state on a subnormal limit cycle crosses -0; noise of ~1e-36 keeps a stream in that zone for good, so that
    nearly every tile is repaired, in every segment position, the last tile of a launch included.  All must leave
    upstream's bits, and the repair must have run."""
    from tests.util import emu_lib
    L = emu_lib()
    if L.ns_emu_hp_spec() == 0:
        pytest.skip("built with -DNS_HP_PAR=0")
    L.ns_emu_hp_respeculated.restype = __import__("ctypes").c_longlong
    hp = L.ns_emu_state_hp_offset()
    n_frames = 14
    rng = np.random.default_rng(11)
    x = np.zeros((5, n_frames * 480), np.float32)
    x[1, :960] = (3000.0 * rng.standard_normal(960)).astype(np.float32)  # noise, then digital silence
    x[3] = (2500.0 * rng.standard_normal(n_frames * 480)).astype(np.float32)  # ordinary input
    x[4] = (1e-36 * rng.standard_normal(n_frames * 480)).astype(np.float32)  # lives where the low product underflows
    state = np.zeros((5, L.ns_emu_state_floats()), np.float32)
    state[0, hp:hp + 2] = (3.1e-36, -2.9e-36)  # decays through 2^-126 within a few thousand samples
    state[1, hp:hp + 2] = (-120.5, 118.25)
    state[2, hp:hp + 2] = np.array([0x80000000 | 249, 497], np.uint32).view(np.float32)  # one step from -0
    start = state[:, hp:hp + 2].copy()
    L.ns_emu_hp_respeculated()
    _, _, _, st = emu_process(model_blob, x, chunk=7, state=state)
    redone = L.ns_emu_hp_respeculated()
    for s in range(5):
        y, m0, m1 = _biquad_reference(x[s], *start[s])
        got = st[s, hp:hp + 2]
        assert got.view(np.uint32).tolist() == [np.float32(m0).view(np.uint32), np.float32(m1).view(np.uint32)], s
        assert np.array_equal(st[s, :1440].view(np.uint32), y[-1440:].view(np.uint32)), s  # the high-passed history
    assert redone > 2 * 7 * 5 * 0.8, redone  # stream 4 alone: nearly each of its 70 tiles
    # ordinary input alone never takes the slow path
    L.ns_emu_hp_respeculated()
    emu_process(model_blob, x[3:4], chunk=7)
    assert L.ns_emu_hp_respeculated() == 0


def test_both_forms_of_the_biquad_kernel_give_the_same_bits(model_blob, sig, monkeypatch):
    """K0 exists as one recursion warp (full batches) and parallel in time (small batches): same output, same state."""
    a = emu_process(model_blob, sig, chunk=8)
    monkeypatch.setenv("CRISPY_NS_HP_PAR", "0")
    b = emu_process(model_blob, sig, chunk=8)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(canonical_state(a[3]).view(np.uint32), canonical_state(b[3]).view(np.uint32))


@pytest.mark.parametrize("hp_par", ["1", "0"])  # both forms of the biquad kernel
def test_non_finite_input_poisons_only_its_own_stream(oracle_model, model_blob, hp_par, monkeypatch):
    """NaN, Inf or near-overflow samples in one recording (a glitching capture device) must not reach its batch
    neighbours: streams share CTAs in every kernel (32 per biquad warp, 16 per recurrent-core CTA whose matrix
    products run one stream per accumulator row).  The neighbours' output, VAD and taps are bit-identical to a run
    without the glitches; the glitched streams go non-finite from the same frame on as the oracle's do (upstream keeps
    no guard either: a DenoiseState that has seen a NaN stays NaN until it is rebuilt, audio.rs:955-965)."""
    monkeypatch.setenv("CRISPY_NS_HP_PAR", hp_par)
    ns, nf = 18, 6  # two recurrent-core groups
    x = make_signal(ns, nf)
    clean = emu_process(model_blob, x, chunk=6)
    bad = x.copy()
    bad[3, 480 * 2 + 17] = np.nan
    bad[17, 480 * 1 + 5] = np.inf
    bad[8, 480 * 3: 480 * 3 + 4] = [3e38, -3e38, 3e38, -3e38]
    dirty = emu_process(model_blob, bad, chunk=6)
    keep = [s for s in range(ns) if s not in (3, 17, 8)]
    for a, b in zip(clean[:3], dirty[:3]):  # output, VAD, taps
        assert np.array_equal(a[keep], b[keep])
    assert np.array_equal(canonical_state(clean[3])[keep], canonical_state(dirty[3])[keep])
    ref, _ = po.process_streams(oracle_model, bad)
    for s in (3, 17, 8):
        got_nan = np.isnan(dirty[0][s].reshape(nf, 480)).any(1)
        want_nan = np.isnan(ref[s].reshape(nf, 480)).any(1)
        assert np.array_equal(got_nan, want_nan), (s, got_nan, want_nan)
        first = int(np.argmax(want_nan))
        assert np.array_equal(dirty[0][s, : first * 480], clean[0][s, : first * 480])  # untouched before the glitch
