"""GPU: the CUDA path through the C ABI (libcrispy_ns.so) against the oracle -- the parity tests proper.
Tolerances are BASELINE.json north_star's: max abs <= 1e-3 of full scale, SNR >= 60 dB, VAD within 1e-3."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import crispy_b200 as cb  # noqa: E402
from crispy_b200.synth import synth_chunk  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tests.util import TOL_MAX_ABS, assert_parity, make_signal, snr_db  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def model():
    return cb.Model.synthetic(0)


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_c1_single_stream_process_frame(oracle_model):
    """configs[0]: one 10 s clip through DenoiseState.new / process_frame (audio.rs:229, :268)."""
    x = make_signal(1, 1000)[0]
    st = cb.DenoiseState.new()
    out = np.zeros_like(x)
    vad = np.zeros(1000, np.float32)
    for t in range(1000):
        vad[t] = st.process_frame(out[t * 480:(t + 1) * 480], x[t * 480:(t + 1) * 480])
    ref, rvad = po.process_streams(oracle_model, x[None, :])
    r = assert_parity(ref[0], out, rvad[0], vad, "C1")
    print("C1 parity", r)
    with pytest.raises(AssertionError):
        st.process_frame(np.zeros(479, np.float32), np.zeros(479, np.float32))


def test_process_frame_graph_replay_equals_plain_launches(model, monkeypatch):
    """crispy_ns_process_frame replays a CUDA graph per workspace slot; CRISPY_NS_NO_GRAPH=1 takes the plain launches."""
    x = make_signal(1, 40)[0].reshape(40, 480)
    outs = []
    for no_graph in (False, True):
        if no_graph:
            monkeypatch.setenv("CRISPY_NS_NO_GRAPH", "1")
        st = cb.DenoiseState.new(model)
        o = np.zeros(480, np.float32)
        frames, vads = [], []
        for t in range(40):
            vads.append(st.process_frame(o, x[t]))
            frames.append(o.copy())
        outs.append((np.concatenate(frames), np.array(vads, np.float32)))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_batch_matches_oracle_with_taps(oracle_model, model):
    n_streams, n_frames = 37, 120
    x = make_signal(n_streams, n_frames)
    den = cb.BatchDenoiser(n_streams, model)
    out, vad, taps = den.process_streams(_dev(x), unit_scale=False, return_taps=True)
    out, vad, taps = out.cpu().numpy(), vad.cpu().numpy(), taps.cpu().numpy()
    ref, rvad = po.process_streams(oracle_model, x, n_threads=8)
    r = assert_parity(ref, out, rvad, vad, "batch")
    print("batch parity", r)
    flips, frames = 0, 0
    for s in (0, 1, 17, 36):
        _, t = po.debug_trace(oracle_model, x[s])
        pi = np.array([a["pitch_index"] for a in t])
        flips += int((pi != taps[s, :, 132].astype(int)).sum())
        frames += len(pi)
        assert np.array_equal(taps[s, :, 133].astype(int), np.array([a["silence"] for a in t]))
        gains = np.array([a["gains"] for a in t])
        assert np.max(np.abs(taps[s, :, 42:64] - gains)) < 5e-4  # band gains; bf16 hi+lo activations on the tensor pipe
        assert np.array_equal(taps[s, :, 130], np.array([a["pitch_gain"] for a in t], dtype=np.float32))
    print(f"pitch decision flips vs oracle: {flips}/{frames}")
    assert flips == 0, "the pitch decision chain is computed bit-exactly (ns_pipe.cuh exactness contract)"


@pytest.mark.parametrize("n_streams", [5, 128, 200])
def test_tcgen05_recurrent_core_matches_oracle_and_mma_core(oracle_model, model, n_streams, monkeypatch):
    """K4 on tcgen05 / tensor memory ($CRISPY_NS_RNN=tc5, ns_rnn_tc5.cuh): same tolerances against the oracle as the
    warp-level core, band gains and VAD within 1e-4 of it, chunked calls with state carry bit-identical to one call
    (ragged stream counts: one partly filled 128-stream CTA, exactly one, two)."""
    n_frames = 72
    x = make_signal(n_streams, n_frames)
    res = {}
    for sel in ("mma", "tc5"):
        monkeypatch.setenv("CRISPY_NS_RNN", sel)
        den = cb.BatchDenoiser(n_streams, model)
        out, vad, taps = den.process_streams(_dev(x), unit_scale=False, return_taps=True)
        res[sel] = (out.cpu().numpy(), vad.cpu().numpy(), taps.cpu().numpy())
        if sel == "tc5":
            den.reset()
            a, va = den.process_streams(_dev(x[:, : 40 * 480]), unit_scale=False)
            b, vb = den.process_streams(_dev(x[:, 40 * 480:]), unit_scale=False)
            assert torch.equal(torch.cat([a, b], 1).cpu(), out.cpu()) and torch.equal(torch.cat([va, vb], 1).cpu(), vad.cpu())
    k = min(n_streams, 6)
    ref, rvad = po.process_streams(oracle_model, x[:k], n_threads=k)
    r = assert_parity(ref, res["tc5"][0][:k], rvad, res["tc5"][1][:k], "tc5")
    print("tcgen05 core parity", r)
    assert np.abs(res["tc5"][1] - res["mma"][1]).max() <= 1e-4
    assert np.abs(res["tc5"][2][:, :, 42:64] - res["mma"][2][:, :, 42:64]).max() <= 5e-4
    assert np.array_equal(res["tc5"][2][:, :, 132], res["mma"][2][:, :, 132])  # pitch index: untouched by the core


def test_golden_fixture(oracle_model, model):
    g = np.load(os.path.join(GOLDEN, "c1_head.npz"))
    x = g["x_i16"][None, :]
    den = cb.BatchDenoiser(1, model)
    out, vad = den.process_streams(_dev(x), unit_scale=False)  # int16 in, f32 (16-bit scale) out
    assert_parity(g["out"][None, :], out.cpu().numpy(), g["vad"][None, :], vad.cpu().numpy(), "golden")


@pytest.mark.parametrize("chunk", [1, 5, 8, 24, 64])
def test_chunk_size_variants_agree(model, chunk, monkeypatch):
    """The engine cuts a call into chunks of CRISPY_NS_CHUNK_FRAMES frames; results must not depend on it."""
    x = _dev(make_signal(19, 30))
    monkeypatch.setenv("CRISPY_NS_CHUNK_FRAMES", "30")
    a, va = cb.BatchDenoiser(19, model).process_streams(x, unit_scale=False)
    monkeypatch.setenv("CRISPY_NS_CHUNK_FRAMES", str(chunk))
    den = cb.BatchDenoiser(19, model)
    assert den.info["chunk_frames"] == chunk
    b, vb = den.process_streams(x, unit_scale=False)
    assert torch.equal(a, b) and torch.equal(va, vb)


def test_synthesis_run_length_does_not_change_results(model, monkeypatch):
    """K5 splits a chunk into runs of frames per task (halo = one re-synthesised frame); any split gives the same bits."""
    x = _dev(make_signal(7, 70))
    base, bv = cb.BatchDenoiser(7, model).process_streams(x, unit_scale=False)
    for run in (1, 3, 8, 1000):
        monkeypatch.setenv("CRISPY_NS_SYN_RUN", str(run))
        out, vad = cb.BatchDenoiser(7, model).process_streams(x, unit_scale=False)
        assert torch.equal(out, base) and torch.equal(vad, bv), run


def test_chunking_save_load_and_reset_are_bit_exact(model):
    x = _dev(make_signal(11, 64))
    den = cb.BatchDenoiser(11, model)
    full, vfull = den.process_streams(x, unit_scale=False)
    den.reset()
    a, va = den.process_streams(x[:, : 20 * 480].contiguous(), unit_scale=False)
    blob = den.save_state()
    b, vb = den.process_streams(x[:, 20 * 480:].contiguous(), unit_scale=False)
    assert torch.equal(torch.cat([a, b], 1), full) and torch.equal(torch.cat([va, vb], 1), vfull)
    other = cb.BatchDenoiser(11, model)
    other.load_state(blob)
    assert other.frames_done == 20
    b2, _ = other.process_streams(x[:, 20 * 480:].contiguous(), unit_scale=False)
    assert torch.equal(b2, b)
    # strided input (rows longer than the processed span)
    den.reset()
    c, _ = den.process_streams(x[:, : 20 * 480], unit_scale=False)
    assert torch.equal(c, a)


def test_wrapper_semantics_unit_scale_volume_drop_first(oracle_model, model):
    xu = (make_signal(5, 40) / 32768.0).astype(np.float32)
    den = cb.BatchDenoiser(5, model)
    out, vad = den.process_streams(_dev(xu), unit_scale=True, volume=0.5, drop_first_frame=True)
    assert out.shape[1] == 39 * 480
    for s in range(5):
        ref = po.processor_run(oracle_model, xu[s], 48000.0, 0.5)  # RnnNoiseProcessor::push_sample
        assert len(ref) == 39 * 480
        assert np.max(np.abs(out[s].cpu().numpy() - ref)) <= 1e-3 and snr_db(ref, out[s].cpu().numpy()) >= 60
    # a second chunk continues without dropping anything
    more, _ = den.process_streams(_dev(xu[:, : 4 * 480]), unit_scale=True, volume=0.5, drop_first_frame=True)
    assert more.shape[1] == 4 * 480


def test_host_path_equals_device_path(model, monkeypatch):
    x = make_signal(13, 90)
    xu = (x / 32768.0).astype(np.float32)
    den = cb.BatchDenoiser(13, model)
    d_out, d_vad = den.process_streams(_dev(xu), unit_scale=True)
    monkeypatch.setenv("CRISPY_NS_HOST_CHUNK_FRAMES", "17")  # force several pipelined chunks
    den.reset()
    hx = torch.from_numpy(xu).pin_memory()
    h_out, h_vad = den.process_streams_host(hx, unit_scale=True)
    assert torch.equal(h_out, d_out.cpu()) and torch.equal(h_vad, d_vad.cpu())
    den.reset()
    h2, _ = den.process_streams_host(hx, unit_scale=True, drop_first_frame=True)
    assert torch.equal(h2, d_out.cpu()[:, 480:])


def test_unaligned_rows_take_the_scalar_loader_and_match(model):
    """K0 fetches 16-byte aligned rows with cp.async (f32 and PCM16) and anything else sample by sample; K5 stores
    16 bytes at a time only into aligned rows: odd row strides and offsets must give the same bits."""
    x = make_signal(5, 40)
    xi = np.clip(np.rint(x), -32768, 32767).astype(np.int16)
    for arr, kw in ((x.astype(np.float32), {}), (xi, {"out_i16": True})):
        ref, rv = cb.BatchDenoiser(5, model).process_streams(_dev(arr), unit_scale=False, **kw)
        wide = torch.zeros((5, arr.shape[1] + 7), dtype=_dev(arr).dtype, device="cuda")
        view = wide[:, 3:3 + arr.shape[1]]
        view.copy_(_dev(arr))
        out_wide = torch.zeros((5, arr.shape[1] + 5), dtype=ref.dtype, device="cuda")
        out_view = out_wide[:, 1:1 + arr.shape[1]]
        out, vad = cb.BatchDenoiser(5, model).process_streams(view, unit_scale=False, out=out_view, **kw)
        assert torch.equal(out, ref) and torch.equal(vad, rv)
        assert float(out_wide[:, 0].abs().max()) == 0 and float(out_wide[:, 1 + arr.shape[1]:].abs().max()) == 0


def test_i16_io_and_dual_mono_mix(oracle_model, model):
    x = make_signal(6, 50)
    xi = np.clip(np.rint(x), -32768, 32767).astype(np.int16)
    den = cb.BatchDenoiser(6, model)
    o16, _ = den.process_streams(_dev(xi), unit_scale=False, out_i16=True)
    ref, _ = po.process_streams(oracle_model, xi.astype(np.float32))
    assert np.max(np.abs(o16.cpu().numpy().astype(np.float64) - np.rint(ref))) <= TOL_MAX_ABS + 1
    # f1: mic denoised + app raw -> clamp -> PCM16 dual mono (commands/recording.rs:260-264)
    xu = (x / 32768.0).astype(np.float32)
    app = (np.random.default_rng(5).standard_normal(xu.shape) * 0.2).astype(np.float32)
    den.reset()
    mix, _ = den.process_streams(_dev(xu), unit_scale=True, app=_dev(app), mix_stereo_i16=True)
    mix = mix.cpu().numpy()
    refd, _ = po.process_streams(oracle_model, xu, unit_scale=True)
    for s in range(6):
        want = po.mix_dual_mono_i16(refd[s], app[s]).reshape(-1, 2)
        assert np.max(np.abs(mix[s].astype(np.int32) - want.astype(np.int32))) <= 34
        assert np.array_equal(mix[s][:, 0], mix[s][:, 1])


def test_linear_resample_is_bit_exact():  # audio.rs:73-134
    rng = np.random.default_rng(7)
    x = rng.standard_normal((3, 44100)).astype(np.float32)
    for rin, rout in ((44100.0, 48000.0), (48000.0, 16000.0), (16000.0, 48000.0), (48000.0, 48000.0)):
        y = cb.linear_resample(_dev(x), rin, rout).cpu().numpy()
        for s in range(3):
            assert np.array_equal(y[s], po.linear_resample(x[s], rin, rout))


def test_config3_441k_front_end(oracle_model, model):
    """configs[2]: 44.1 kHz input, linear resample to 48 kHz ahead of the denoiser (audio.rs:217-221)."""
    x44 = (synth_chunk(4, 44100 * 2).numpy()).astype(np.float32)
    y = cb.linear_resample(_dev(x44), 44100.0, 48000.0)
    nfr = y.shape[1] // 480
    den = cb.BatchDenoiser(4, model)
    out, _ = den.process_streams(y[:, : nfr * 480], unit_scale=True, drop_first_frame=True)
    for s in range(4):
        ref = po.processor_run(oracle_model, x44[s], 44100.0, 1.0)
        assert len(ref) == out.shape[1]
        assert np.max(np.abs(out[s].cpu().numpy() - ref)) <= 1e-3


def test_sinc_resample_is_bit_exact():  # north_star item 4; oracle: rno_sinc_resample
    rng = np.random.default_rng(11)
    x = rng.standard_normal((3, 44100 + 17)).astype(np.float32)
    for rin, rout in ((44100, 48000), (48000, 44100), (16000, 48000), (48000, 16000), (22050, 48000), (48000, 8000)):
        y = cb.sinc_resample(_dev(x), rin, rout).cpu().numpy()
        for s in range(3):
            want = po.sinc_resample(x[s], rin, rout)
            assert y.shape[1] == len(want)
            assert np.array_equal(y[s], want), (rin, rout, np.abs(y[s] - want).max())
    # ragged edges: shorter than the filter, a single sample, other filter lengths, strided rows
    for n_in in (1, 5, 127, 441, 442):
        y = cb.sinc_resample(_dev(x[:, :n_in]), 44100, 48000).cpu().numpy()
        assert np.array_equal(y[1], po.sinc_resample(x[1, :n_in], 44100, 48000))
    y = cb.sinc_resample(_dev(x)[:, 100:5000], 44100, 48000, sinc_len=64, f_cutoff=0.9).cpu().numpy()
    assert np.array_equal(y[2], po.sinc_resample(x[2, 100:5000], 44100, 48000, 64, 0.9))


def test_sinc_resample_chunks_reproduce_the_whole_call():
    """A recording too long for one call: outputs in chunks of whole periods from the input window their taps touch."""
    rng = np.random.default_rng(4)
    x = rng.standard_normal((3, 44100)).astype(np.float32)
    whole = cb.sinc_resample(_dev(x), 44100, 48000)
    n_total, n_out_total = x.shape[1], whole.shape[1]
    pieces, first = [], 0
    for n_out in (480 * 7, 160, 480 * 40, n_out_total):  # the last one is clipped to what is left
        n_out = min(n_out, n_out_total - first)
        lo, n_in = cb.sinc_needed(44100, 48000, n_total, first, n_out)
        pieces.append(cb.sinc_resample_chunk(_dev(x[:, lo:lo + n_in]), lo, n_total, first, n_out, 44100, 48000))
        first += n_out
    assert first == n_out_total
    assert torch.equal(torch.cat(pieces, 1), whole)
    with pytest.raises(cb.CrispyNsError):  # window too short for the taps
        cb.sinc_resample_chunk(_dev(x[:, 5000:6000]), 5000, n_total, 4800, 4800, 44100, 48000)


def test_front_end_golden_fixture():
    """The committed front-end fixture (tests/golden/front_end.npz) through the CUDA kernels: both bit-exact."""
    g = np.load(os.path.join(GOLDEN, "front_end.npz"))
    x = _dev(g["x44"][None, :])
    assert np.array_equal(cb.sinc_resample(x, 44100, 48000)[0].cpu().numpy(), g["sinc"])
    assert np.array_equal(cb.linear_resample(x, 44100.0, 48000.0)[0].cpu().numpy(), g["linear"])


def test_resample_host_matches_device_paths():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((3, 4410)).astype(np.float32)
    assert np.array_equal(cb.resample_host(x, 44100, 48000, "sinc"), cb.sinc_resample(_dev(x), 44100, 48000).cpu().numpy())
    assert np.array_equal(cb.resample_host(x, 44100, 48000, "linear"), cb.linear_resample(_dev(x), 44100.0, 48000.0).cpu().numpy())


def test_sinc_resample_properties_at_full_size():
    """configs[2] geometry (1,024 streams x 10 s at 44.1 kHz): linearity in the input, exactly one frame
    per 441 samples, and time-shift invariance by whole periods (147 in -> 160 out)."""
    n_streams, n_in = 1024, 441000
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn((n_streams, n_in), device="cuda", generator=g)
    b = torch.randn((n_streams, n_in), device="cuda", generator=g)
    ya, yb, yab = (cb.sinc_resample(t, 44100, 48000) for t in (a, b, a + b))
    assert ya.shape == (n_streams, 480000)
    assert float((yab - (ya + yb)).abs().max()) < 2e-5
    shifted = cb.sinc_resample(a[:, 147 * 5:], 44100, 48000)
    assert torch.equal(shifted[:, 200:-200], ya[:, 160 * 5 + 200:-200])
    # a 1 kHz tone comes out as a 1 kHz tone
    t = torch.arange(n_in, device="cuda", dtype=torch.float64) / 44100.0
    tone = (0.5 * torch.sin(2 * np.pi * 1000.0 * t)).float()[None, :].contiguous()
    yt = cb.sinc_resample(tone, 44100, 48000)[0].double()
    want = 0.5 * torch.sin(2 * np.pi * 1000.0 * torch.arange(480000, device="cuda", dtype=torch.float64) / 48000.0)
    assert float((yt - want)[400:-400].abs().max()) < 2e-5


def test_config3_441k_sinc_front_end(oracle_model, model):
    """configs[2] as north_star words it: 44.1 kHz input, sinc resample to 48 kHz ahead of the analysis."""
    x44 = (synth_chunk(4, 44100 * 2).numpy()).astype(np.float32)
    den = cb.BatchDenoiser(4, model)
    out, vad = den.process_streams(_dev(x44), unit_scale=True, input_rate=44100, front_end="sinc")
    assert out.shape[1] == 200 * 480
    for s in range(4):
        y48 = po.sinc_resample(x44[s], 44100, 48000)
        ref, rvad = po.process_streams(oracle_model, y48[None, :], unit_scale=True)
        o = out[s].cpu().numpy()
        assert np.max(np.abs(o - ref[0])) <= 1e-3 and snr_db(ref[0], o) >= 60, f"sinc front end, stream {s}"
        assert np.max(np.abs(vad[s].cpu().numpy() - rvad[0])) <= 1e-3


def test_full_size_properties(model):
    """configs[1] geometry (1,024 streams) on a 6 s slice: results must not depend on the
    engine's chunk size or on how the caller splits the recording, silence must reconstruct exactly, and nothing may be NaN."""
    n_streams, n_frames = 1024, 600
    x = torch.cat([synth_chunk(n_streams, 100 * 480, start_sample=f * 480, device="cuda") for f in range(0, n_frames, 100)], 1)
    den = cb.BatchDenoiser(n_streams, model)
    out, vad = den.process_streams(x, unit_scale=True)
    assert torch.isfinite(out).all() and torch.isfinite(vad).all()
    assert float(vad.min()) >= 0.0 and float(vad.max()) <= 1.0
    den.reset()
    o1, _ = den.process_streams(x[:, : 250 * 480], unit_scale=True)
    o2, _ = den.process_streams(x[:, 250 * 480:], unit_scale=True)
    assert torch.equal(torch.cat([o1, o2], 1), out)
    os.environ["CRISPY_NS_CHUNK_FRAMES"] = "56"
    try:
        alt, _ = cb.BatchDenoiser(n_streams, model).process_streams(x, unit_scale=True)
    finally:
        del os.environ["CRISPY_NS_CHUNK_FRAMES"]
    assert torch.equal(alt, out)
    # muted stretches (stream % 16 == 3, second half of every 4 s) come out as exact zeros after ring-out
    assert float(out[3, 350 * 480:400 * 480].abs().max()) < 1e-3


def test_edge_cases_empty_single_frame_wide_batch_and_bad_arguments(oracle_model, model):
    """Empty calls leave every state untouched; one frame of one stream; a batch far wider than the SM count (4,100
    streams: partly filled 16- and 32-stream groups at the end) against the oracle on its first, a middle and its last
    streams; DROP_FIRST_FRAME applies to the first call only (audio.rs:275-278); bad arguments come back as error
    codes, never as exceptions through the C ABI."""
    den = cb.BatchDenoiser(3, model)
    x = _dev(make_signal(3, 20))
    o0, v0 = den.process_streams(x[:, :0].contiguous(), unit_scale=False)
    assert o0.shape == (3, 0) and v0.shape == (3, 0) and den.frames_done == 0
    full, vfull = den.process_streams(x, unit_scale=False)
    den.reset()
    a, va = den.process_streams(x[:, :480].contiguous(), unit_scale=False)     # a single frame
    den.process_streams(x[:, :0].contiguous(), unit_scale=False)               # an empty call in between
    b, vb = den.process_streams(x[:, 480:].contiguous(), unit_scale=False)
    assert torch.equal(torch.cat([a, b], 1), full) and torch.equal(torch.cat([va, vb], 1), vfull)
    one = cb.BatchDenoiser(1, model)
    xs = make_signal(1, 1)
    o1, v1 = one.process_streams(_dev(xs), unit_scale=False)
    ref, rvad = po.process_streams(oracle_model, xs)
    assert_parity(ref, o1.cpu().numpy(), rvad, v1.cpu().numpy(), "one frame")
    # drop_first_frame: the first call returns one frame less, later calls do not
    den.reset()
    d1, _ = den.process_streams(x[:, : 5 * 480].contiguous(), unit_scale=False, drop_first_frame=True)
    d2, _ = den.process_streams(x[:, 5 * 480:].contiguous(), unit_scale=False, drop_first_frame=True)
    assert d1.shape[1] == 4 * 480 and torch.equal(torch.cat([d1, d2], 1), full[:, 480:])
    # wide batch
    n_wide, nf = 4100, 12
    xw = synth_chunk(n_wide, nf * 480, device="cuda")
    wide = cb.BatchDenoiser(n_wide, model)
    ow, vw = wide.process_streams(xw, unit_scale=True)
    ids = [0, 1, 2047, 4095, 4096, 4099]
    refw, rvw = po.process_streams(oracle_model, xw[ids].cpu().numpy(), unit_scale=True, n_threads=6)
    err = np.abs(ow[ids].cpu().numpy().astype(np.float64) - refw).max()
    assert err <= 1e-3 and np.abs(vw[ids].cpu().numpy() - rvw).max() <= 1e-3, err
    # bad arguments
    from crispy_b200 import _lib
    L = _lib.lib()
    assert L.crispy_ns_process_streams(None, None, None, None, None, 1, 480, 480, 1, 0, 0, 1.0, None) != 0
    assert b"bad argument" in L.crispy_ns_last_error()
    with pytest.raises(cb.CrispyNsError):
        den.process_streams(x[:, :480].contiguous(), unit_scale=False, out_i16=True, volume=0.5)


def test_resample_audio_is_bit_exact():  # recording.rs:13-39: the recorder's app-audio resampler
    rng = np.random.default_rng(11)
    for n, fr, to in ((1, 44100, 48000), (2, 44100, 48000), (441, 44100, 48000), (44101, 44100, 48000), (4800, 48000, 44100),
                      (1000, 16000, 48000), (999, 48000, 16000), (777, 48000, 48000)):
        x = rng.standard_normal((3, n)).astype(np.float32)
        got = cb.resample_audio(_dev(x), fr, to).cpu().numpy()
        for s in range(3):
            want = po.resample_audio(x[s], fr, to)
            assert got[s].shape == want.shape and np.array_equal(got[s], want), (n, fr, to, s)
    # rows longer than the resampled span, odd element offset (strided, unaligned input)
    big = torch.from_numpy(rng.standard_normal((4, 50003)).astype(np.float32)).cuda()
    view = big[:, 1:44102]
    got = cb.resample_audio(view, 44100, 48000).cpu().numpy()
    assert np.array_equal(got[2], po.resample_audio(view[2].cpu().numpy(), 44100, 48000))
    # the host-pointer entry point (what the Rust shim calls): kind 2
    xh = rng.standard_normal((3, 4411)).astype(np.float32)
    yh = cb.resample_host(xh, 44100, 48000, kind="audio")
    assert np.array_equal(yh[1], po.resample_audio(xh[1], 44100, 48000))
    # configs[3] with a 44.1 kHz app source: resample_audio -> the dual-mono mix beside the denoised microphone
    nf = 20
    mic = synth_chunk(2, nf * 480, device="cuda")
    app44 = synth_chunk(2, nf * 441, first_stream=9, device="cuda") * 0.5
    app = cb.resample_audio(app44, 44100, 48000)
    assert app.shape[1] == nf * 480
    den = cb.BatchDenoiser(2, cb.Model.synthetic(0))
    mix, _ = den.process_streams((mic * 32767.0).round().clamp(-32768, 32767).to(torch.int16), unit_scale=True, app=app, mix_stereo_i16=True)
    assert mix.shape == (2, nf * 480, 2) and torch.equal(mix[:, :, 0], mix[:, :, 1])


def test_downmix_mono_is_bit_exact():  # audio.rs:754-755, :816-818, :879-884: what the capture callbacks hand to push_sample
    rng = np.random.default_rng(5)
    n = 1003
    for ch in (1, 2, 3, 6):
        for dt in (np.float32, np.int16, np.uint16):
            if dt == np.float32:
                x = rng.standard_normal((3, n * ch)).astype(np.float32)
            elif dt == np.int16:
                x = rng.integers(-32768, 32768, (3, n * ch)).astype(np.int16)
            else:
                x = rng.integers(0, 65536, (3, n * ch)).astype(np.uint16)
            got = cb.downmix_mono(torch.from_numpy(x).cuda(), ch).cpu().numpy()
            for s in range(3):
                assert np.array_equal(got[s], po.downmix_mono(x[s], ch)), (ch, dt, s)
    # a stereo row that starts on an odd element: the scalar path
    x = rng.standard_normal((2, 2 * n + 1)).astype(np.float32)
    got = cb.downmix_mono(torch.from_numpy(x).cuda()[:, 1:], 2).cpu().numpy()
    assert np.array_equal(got[1], po.downmix_mono(x[1, 1:], 2))
    # stereo PCM16 capture -> mono -> denoise: the batched form of the i16 callback feeding push_sample
    nf = 12
    st = (synth_chunk(2, nf * 480, device="cuda") * 32767.0).round().clamp(-32768, 32767).to(torch.int16)
    stereo = torch.stack([st, st], 2).reshape(2, -1).contiguous()
    mono = cb.downmix_mono(stereo, 2)
    assert torch.equal(mono, (st.to(torch.float32) / 32768.0 + st.to(torch.float32) / 32768.0) / 2.0)
    out, _ = cb.BatchDenoiser(2, cb.Model.synthetic(0)).process_streams(mono, unit_scale=True)
    assert out.shape == mono.shape and bool(torch.isfinite(out).all())


def test_both_sinc_kernels_give_the_same_bits(monkeypatch):
    """44.1 -> 48 kHz takes the second-generation kernel (four adjacent outputs per thread); CRISPY_NS_SINC_V1=1 forces
    the first one.  Both accumulate fmaf(h[k], x[start + k], acc) in ascending k, so they agree bit for bit -- also on
    rows that start on an odd element (scalar stores) and on a partial last tile."""
    rng = np.random.default_rng(13)
    x = torch.from_numpy(rng.standard_normal((5, 44100 * 2 + 17)).astype(np.float32)).cuda()
    for view in (x, x[:, 1:4412], x[:, :441]):
        b = cb.sinc_resample(view, 44100, 48000)
        monkeypatch.setenv("CRISPY_NS_SINC_V1", "1")
        a = cb.sinc_resample(view, 44100, 48000)
        monkeypatch.delenv("CRISPY_NS_SINC_V1")
        assert torch.equal(a, b)
    want = po.sinc_resample(x[0, :4410].cpu().numpy(), 44100, 48000)
    assert np.array_equal(cb.sinc_resample(x[:1, :4410].contiguous(), 44100, 48000)[0].cpu().numpy(), want)
    # 32 kHz -> 48 kHz sits on the edge of the second kernel's range (M / L = 2 / 3)
    y = cb.sinc_resample(x[:2, :32000].contiguous(), 32000, 48000)
    assert np.array_equal(y[1].cpu().numpy(), po.sinc_resample(x[1, :32000].cpu().numpy(), 32000, 48000))


def test_both_forms_of_the_biquad_kernel_give_the_same_bits(model, monkeypatch):
    """K0 runs as one recursion warp beside the pitch CTAs of a full batch and parallel in time (one warp speculating in
    f32, four running upstream's f64 expression from its recorded states) on small batches; CRISPY_NS_HP_PAR forces
    either.  Same output, VAD and state bit for bit -- also through digital silence, where the speculation is repaired
    (a state decaying through 2^-126), on PCM16 input and with a ragged last CTA."""
    x = make_signal(70, 120, seed=77)
    x[3, 480 * 4:] = 0.0          # decays into digital silence: the low product underflows on the way
    x[40] *= np.float32(3e-40)    # ~1e-36: lives where the speculation misses all the time
    x[69, : 480 * 50] = 0.0
    xi = np.clip(np.rint(x), -32768, 32767).astype(np.int16)
    res = []
    for par in ("1", "0"):
        monkeypatch.setenv("CRISPY_NS_HP_PAR", par)
        den = cb.BatchDenoiser(70, model)
        out, vad = den.process_streams(_dev(x), unit_scale=False)
        st = den.save_state()
        den16 = cb.BatchDenoiser(70, model)
        o16, _ = den16.process_streams(_dev(xi), unit_scale=False, out_i16=True)
        res.append((out.cpu(), vad.cpu(), st, o16.cpu()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    assert res[0][2] == res[1][2]
    assert torch.equal(res[0][3], res[1][3])
