"""Where does the GPU pitch gain differ in bits from the oracle's?  (1,024 x 6,000 frames run, 16 streams compared)"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import crispy_b200 as cb
from crispy_b200.synth import synth_chunk
from oracle import pyoracle as po

ids = [0, 3, 19, 64, 127, 200, 255, 256, 333, 511, 512, 640, 777, 900, 1003, 1023]
n_frames = 6000
x = torch.cat([synth_chunk(1024, 100 * 480, start_sample=f * 480, device="cuda") for f in range(0, n_frames, 100)], 1)[ids].contiguous()
model = cb.Model.synthetic(0)
den = cb.BatchDenoiser(16, model)
o, v, taps = den.process_streams(x, unit_scale=True, return_taps=True)
taps = taps.cpu().numpy()
om = po.Model.synthetic(0)
ref, rvad, rpi, rpg, rsil = po.process_streams_trace(om, x.cpu().numpy(), unit_scale=True, n_threads=16, native=True)
bad = np.argwhere(taps[:, :, 130] != rpg)
print(len(bad), "mismatches")
for s, t in bad[:40]:
    g, r = taps[s, t, 130], rpg[s, t]
    print(f"stream {ids[s]} frame {t}: gpu {g!r} ({g.view(np.uint32):08x}) oracle {r!r} ({r.view(np.uint32):08x}) pitch {int(taps[s,t,132])}/{rpi[s,t]} sil {int(taps[s,t,133])}/{rsil[s,t]}")
# also against the -O2 checker build of the oracle
ref2, _, rpi2, rpg2, _ = po.process_streams_trace(om, x.cpu().numpy(), unit_scale=True, n_threads=16, native=False)
print("native vs O2 oracle: gain diffs", int((rpg2 != rpg).sum()), "pitch diffs", int((rpi2 != rpi).sum()), "out equal", np.array_equal(ref, ref2))
print("gpu vs O2 oracle gain diffs", int((taps[:, :, 130] != rpg2).sum()))
