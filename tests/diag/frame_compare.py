"""GPU vs oracle, band by band, at given frames of one synthetic stream (state carried from frame 0):
the batch is synthesised exactly as tests/test_gpu_configs.py does (first_stream, n_streams, pieces of 100 frames: the
noise draw depends on all three).
usage: python tests/diag/frame_compare.py <first_stream> <n_streams> <index in the batch> <frame> [more frames ...]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import crispy_b200 as cb  # noqa: E402
from crispy_b200.synth import synth_chunk  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

first, n, idx = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
frames = sorted(int(a) for a in sys.argv[4:])
nf = (frames[-1] + 2 + 99) // 100 * 100
x = torch.cat([synth_chunk(n, 100 * 480, first_stream=first, start_sample=f * 480, device="cuda") for f in range(0, nf, 100)], 1)
den = cb.BatchDenoiser(n, cb.Model.synthetic(0))
outs, taps = [], []
for f in range(0, nf, 6000):  # the same 60 s calls as the long-run test
    o, v, tp = den.process_streams(x[:, f * 480:min(nf, f + 6000) * 480].contiguous(), unit_scale=True, return_taps=True)
    outs.append(o.cpu().numpy()), taps.append(tp.cpu().numpy())
o, taps = np.concatenate(outs, 1)[idx], np.concatenate(taps, 1)[idx]
xs = x.cpu().numpy()[idx]
om = po.Model.synthetic(0)
st = po.DenoiseState(om)
L = po.lib()
L.rno_get_raw_gains.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
want = set()
for f in frames:
    want.update((f - 1, f, f + 1))
for t in range(nf):
    ro, rv = st.process_frame((xs[t * 480:(t + 1) * 480] * 32768.0).astype(np.float32))
    if t not in want:
        continue
    d = st.debug()
    graw = (C.c_float * 22)()
    L.rno_get_raw_gains(st.h, graw)
    graw = np.array(graw[:], np.float32)
    err = np.abs(o[t * 480:(t + 1) * 480].astype(np.float64) - ro / 32768.0)
    T = taps[t]
    print(f"frame {t}: max err {err.max():.3e} FS at sample {err.argmax()}; pitch gpu {int(T[132])} ref {d['pitch_index']}; "
          f"max|feat| diff {np.abs(T[:42] - d['features']).max():.2e}; vad gpu {T[131]:.6f} ref {rv:.6f}")
    for b in range(22):
        print(f"   band {b:2d} Exp {T[108 + b]: .7f} / {d['Exp'][b]: .7f}  graw {T[136 + b]:.7f} / {graw[b]:.7f}  g {T[42 + b]:.7f} / {d['gains'][b]:.7f}"
              f"  Ex {T[64 + b]:.4e} / {d['Ex'][b]:.4e}  Ep {T[86 + b]:.4e} / {d['Ep'][b]:.4e}"
              + ("   <-- branch differs" if (T[108 + b] > T[136 + b]) != (d['Exp'][b] > graw[b]) else ""))
