"""Property tests (hypothesis) of the host-side logic: stream partition, front-end geometry."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

import crispy_b200 as cb
from crispy_b200 import _lib
from crispy_b200.shard import job_rate, parse_cpulist, stream_block
from oracle import pyoracle as po


@settings(max_examples=200, deadline=None)
@given(units=st.integers(0, 5000), world=st.integers(1, 16), group=st.sampled_from([1, 2, 4]))
def test_stream_blocks_partition_exactly(units, world, group):
    n = units * group
    blocks = [stream_block(n, world, r, group) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n
    for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in blocks]
    assert all(s % group == 0 for s in sizes) and max(sizes) - min(sizes) <= group  # sources of a meeting stay together


@settings(max_examples=100, deadline=None)
@given(rates=st.lists(st.floats(0.1, 1e6), min_size=1, max_size=8))
def test_job_rate_is_total_units_over_slowest_rank(rates):
    units = [1000.0] * len(rates)
    secs = [u / r for u, r in zip(units, rates)]
    assert np.isclose(job_rate(units, secs), sum(units) / max(secs))


@settings(max_examples=100, deadline=None)
@given(cpus=st.sets(st.integers(0, 255), min_size=1, max_size=40))
def test_cpulist_roundtrip(cpus):
    cpus = sorted(cpus)
    parts, i = [], 0
    while i < len(cpus):  # compress to the kernel's "a-b,c" form
        j = i
        while j + 1 < len(cpus) and cpus[j + 1] == cpus[j] + 1:
            j += 1
        parts.append(f"{cpus[i]}-{cpus[j]}" if j > i else f"{cpus[i]}")
        i = j + 1
    assert parse_cpulist(",".join(parts)) == cpus


@settings(max_examples=150, deadline=None)
@given(n_total=st.integers(1, 200000), first_periods=st.integers(0, 1000), n_out=st.integers(1, 5000),
       pair=st.sampled_from([(44100, 48000), (48000, 44100), (16000, 48000), (48000, 16000), (22050, 48000)]))
def test_sinc_window_covers_exactly_the_taps(n_total, first_periods, n_out, pair):
    rin, rout = pair
    g = np.gcd(rin, rout)
    L, M = rout // g, rin // g
    total_out = _lib.lib().crispy_ns_sinc_resample_count(rin, rout, n_total)
    assert total_out == -(-n_total * L // M) == po.lib().rno_sinc_resample_count(rin, rout, n_total)
    first = first_periods * L
    if first >= total_out:
        return
    n_out = min(n_out, total_out - first)
    lo, n_in = cb.sinc_needed(rin, rout, n_total, first, n_out)
    # every tap index of every requested output that lies inside the recording lies inside [lo, lo + n_in)
    base_first, base_last = first * M // L, (first + n_out - 1) * M // L
    want_lo, want_hi = max(0, base_first - 127), min(n_total, base_last + 129)
    assert lo == want_lo and lo + n_in == max(want_hi, want_lo)


@settings(max_examples=30, deadline=None)
@given(n=st.integers(0, 3000), seed=st.integers(0, 2**31 - 1))
def test_sinc_oracle_is_linear_and_bounded(n, seed):
    rng = np.random.default_rng(seed)
    a, b = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    ya, yb, yab = (po.sinc_resample(v, 44100, 48000) for v in (a, b, a + b))
    assert len(ya) == -(-n * 160 // 147)
    if n:
        assert np.abs(yab - (ya + yb)).max() < 5e-5 * max(1.0, np.abs(yab).max())


@settings(max_examples=150, deadline=None)
@given(n=st.integers(0, 5000), rates=st.sampled_from([(44100, 48000), (48000, 44100), (16000, 48000), (48000, 16000), (22050, 48000),
                                                      (8000, 48000), (48000, 48000), (47999, 48000), (96000, 48000)]))
def test_resample_audio_count_and_shape(n, rates):  # recording.rs:19, :27-35
    fr, to = rates
    x = np.linspace(-1.0, 1.0, n, dtype=np.float32)
    y = po.resample_audio(x, fr, to)
    assert _lib.lib().crispy_ns_resample_audio_count(n, fr, to) == len(y)
    if n:
        assert len(y) == int(np.ceil(np.float64(n) / (np.float64(fr) / np.float64(to)))) or fr == to
        assert y[0] == x[0] and np.all(np.diff(y.astype(np.float64)) >= -1e-6)  # an increasing ramp stays one
        assert y.min() >= x.min() and y.max() <= x.max()  # interpolation never leaves the range of its two neighbours


@settings(max_examples=100, deadline=None)
@given(ch=st.integers(1, 8), n=st.integers(0, 64), seed=st.integers(0, 2 ** 31 - 1), fmt=st.sampled_from(["f32", "i16", "u16"]))
def test_downmix_of_identical_channels_is_the_channel(ch, n, seed, fmt):  # audio.rs:754-755, :816-818, :879-884
    rng = np.random.default_rng(seed)
    if fmt == "f32":
        one = rng.standard_normal(n).astype(np.float32)
        unit = one
    elif fmt == "i16":
        one = rng.integers(-32768, 32768, n).astype(np.int16)
        unit = one.astype(np.float32) / np.float32(32768.0)
    else:
        one = rng.integers(0, 65536, n).astype(np.uint16)
        unit = (one.astype(np.float32) - np.float32(32768.0)) / np.float32(32768.0)
    y = po.downmix_mono(np.repeat(one, ch), ch)
    assert y.shape == (n,)
    # exact where neither the running sum nor the division rounds: one or two channels; 16-bit samples (their sums fit
    # the float32 significand) with a power-of-two channel count
    if ch <= 2 or (fmt != "f32" and ch in (4, 8)):
        assert np.array_equal(y, unit)
    else:
        assert np.allclose(y, unit, rtol=3e-7, atol=1e-12)


@settings(max_examples=150, deadline=None)
@given(n=st.integers(0, 6000),
       rin=st.one_of(st.sampled_from([8000.0, 11025.0, 16000.0, 22050.0, 32000.0, 44100.0, 47999.5, 48000.0, 48000.9, 88200.0,
                                      96000.0, 192000.0]), st.floats(4000.0, 200000.0, width=32)),
       rout=st.sampled_from([48000.0, 44100.0, 16000.0]))
def test_linear_resampler_count_matches_the_sample_by_sample_walk(n, rin, rout):  # audio.rs:108-133
    """LinearResampler emits while next_out_pos <= in_pos with f64 positions: the library's closed count (what sizes
    the device output row) equals the length of the oracle's sample-by-sample walk for any device rate, including
    the |in - out| < 1 passthrough (audio.rs:109-112), and the outputs stay inside the range of the input."""
    x = np.linspace(-1.0, 1.0, n, dtype=np.float32)
    y = po.linear_resample(x, rin, rout)
    assert _lib.lib().crispy_ns_linear_resample_count(rin, rout, n) == len(y)
    if abs(np.float32(rin) - np.float32(rout)) < 1.0:
        assert np.array_equal(y, x)
    elif len(y):
        assert y.min() >= x.min() - 1e-6 and y.max() <= x.max() + 1e-6
        assert np.all(np.diff(y.astype(np.float64)) >= -1e-6)
