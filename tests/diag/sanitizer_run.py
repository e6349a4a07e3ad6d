"""A small pass over every kernel of the library for compute-sanitizer (memcheck / racecheck):
  compute-sanitizer --tool memcheck python tests/diag/sanitizer_run.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import crispy_b200 as cb  # noqa: E402
from crispy_b200.synth import synth_chunk  # noqa: E402

n, nf = int(sys.argv[1]) if len(sys.argv) > 1 else 37, int(sys.argv[2]) if len(sys.argv) > 2 else 40
x = synth_chunk(n, nf * 480, device="cuda")
model = cb.Model.synthetic(0)
for sel in ("mma", "tc5"):
    os.environ["CRISPY_NS_RNN"] = sel
    den = cb.BatchDenoiser(n, model)
    o, v, taps = den.process_streams(x, unit_scale=True, return_taps=True)
    o2, v2 = den.process_streams(x[:, : 7 * 480].contiguous(), unit_scale=True, drop_first_frame=True)
    xi = (x * 32767.0).round().clamp(-32768, 32767).to(torch.int16)
    app = torch.roll(x, 1, 0) * 0.5
    den.reset()
    mix, _ = den.process_streams(xi, unit_scale=True, app=app, mix_stereo_i16=True)
    oi, _ = den.process_streams(xi, unit_scale=True, out_i16=True)
    xu = torch.zeros((n, nf * 480 + 3), dtype=torch.float32, device="cuda")[:, 1:nf * 480 + 1]  # unaligned rows
    xu.copy_(x)
    den.reset()
    den.process_streams(xu, unit_scale=True)
    torch.cuda.synchronize()
    print(sel, "ok", float(o.abs().max()), int(mix.abs().max()))
os.environ.pop("CRISPY_NS_RNN")
x44 = synth_chunk(5, 44100 // 10 * 3, device="cuda")
y = cb.sinc_resample(x44, 44100, 48000)
os.environ["CRISPY_NS_SINC_V1"] = "1"
y1 = cb.sinc_resample(x44, 44100, 48000)
os.environ.pop("CRISPY_NS_SINC_V1")
assert torch.equal(y, y1)
y3 = cb.sinc_resample(x44[:, 1:4412], 44100, 48000)  # odd rows: scalar stores
y2 = cb.linear_resample(x44, 44100.0, 48000.0)
den = cb.BatchDenoiser(5, model)
den.process_streams(x44[:, : 441 * 20].contiguous(), unit_scale=True, input_rate=44100, front_end="sinc")
ya = cb.resample_audio(x44, 44100, 48000)
ya2 = cb.resample_audio(x44[:, 1:4412], 44100, 48000)   # strided rows, scalar stores
st16 = (torch.randn((3, 2 * 4001), device="cuda") * 8000).to(torch.int16)
for ch, t in ((2, x44[:, :8000]), (2, st16[:, :8000]), (2, st16[:, 1:]), (3, x44[:, :9000]), (1, x44[:, :77])):
    cb.downmix_mono(t, ch)
hx = x.cpu().pin_memory()
den = cb.BatchDenoiser(n, model)
ho, hv = den.process_streams_host(hx, unit_scale=True)
st = cb.DenoiseState.new(model)
out = np.zeros(480, np.float32)
for t in range(4):
    st.process_frame(out, (hx[0, t * 480:(t + 1) * 480].numpy() * 32768.0).astype(np.float32))
torch.cuda.synchronize()
print("front ends, host path, single frame ok", tuple(y.shape), tuple(y2.shape))
