"""Which frames exceed 1e-3 FS against the oracle on a long run, and why (band gains g vs band correlation Exp)?"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import crispy_b200 as cb
from crispy_b200.synth import synth_chunk
from oracle import pyoracle as po

first, n, n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40, 4, int(sys.argv[2]) if len(sys.argv) > 2 else 24000
x = torch.cat([synth_chunk(n, 100 * 480, first_stream=first, start_sample=f * 480, device="cuda") for f in range(0, n_frames, 100)], 1)
if len(sys.argv) > 3 and sys.argv[3] == "i16":
    x = (x * 32767.0).round().clamp_(-32768, 32767) / 32768.0
model = cb.Model.synthetic(0)
den = cb.BatchDenoiser(n, model)
o, v, taps = den.process_streams(x.contiguous(), unit_scale=True, return_taps=True)
o, taps = o.cpu().numpy(), taps.cpu().numpy()
om = po.Model.synthetic(0)
xs = x.cpu().numpy()
ref, rvad, rpi, rpg, rsil = po.process_streams_trace(om, xs, unit_scale=True, n_threads=n, native=True)
err = np.abs(o - ref).reshape(n, n_frames, 480).max(2)
for s in range(n):
    e = o[s].astype(np.float64) - ref[s]
    print(f"stream {first+s}: max {err[s].max():.2e} frames>1e-3: {(err[s] > 1e-3).sum()} frames>1e-4: {(err[s] > 1e-4).sum()} snr {10*np.log10((ref[s].astype(np.float64)**2).mean()/max((e**2).mean(),1e-30)):.1f} dB")
s, t = np.unravel_index(np.argmax(err), err.shape)
print("worst: stream", first + s, "frame", t, "err", err[s, t])
# oracle taps for that stream up to frame t+2
st = po.DenoiseState(om)
for f in range(t + 2):
    st.process_frame((xs[s, f * 480:(f + 1) * 480] * 32768.0).astype(np.float32))
    if f >= t - 2:
        d = st.debug()
        gg, ge = taps[s, f, 42:64], taps[s, f, 108:130]
        print(f"frame {f}: err {err[s, f]:.2e} max|g_gpu-g_ref| {np.abs(gg - d['gains']).max():.2e} max|Exp_gpu-Exp_ref| {np.abs(ge - d['Exp']).max():.2e}")
        close = np.abs(d['Exp'] - d['gains'])
        print("   bands where |Exp - g| < 1e-3:", [(int(b), float(d['Exp'][b]), float(d['gains'][b]), float(ge[b]), float(gg[b])) for b in np.where(close < 1e-3)[0]])
