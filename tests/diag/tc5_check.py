"""tcgen05 recurrent core ($CRISPY_NS_RNN=tc5) against the mma.sync core and the oracle."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import crispy_b200 as cb
from crispy_b200.synth import synth_chunk
from oracle import pyoracle as po

n, nf = int(sys.argv[1]) if len(sys.argv) > 1 else 200, int(sys.argv[2]) if len(sys.argv) > 2 else 96
x = synth_chunk(n, nf * 480, device="cuda")
outs = {}
for sel in ("mma", "tc5"):
    os.environ["CRISPY_NS_RNN"] = sel
    den = cb.BatchDenoiser(n)
    o, v, taps = den.process_streams(x, unit_scale=True, return_taps=True)
    torch.cuda.synchronize()
    outs[sel] = (o.cpu().numpy(), v.cpu().numpy(), taps.cpu().numpy())
    print(sel, "frames_done", den.frames_done)
a, b = outs["mma"], outs["tc5"]
print("tc5 vs mma: out max", np.abs(a[0] - b[0]).max(), "vad max", np.abs(a[1] - b[1]).max(), "gains max", np.abs(a[2][:, :, 42:64] - b[2][:, :, 42:64]).max())
k = min(n, 8)
ref, rv = po.process_streams(po.Model.synthetic(0), x[:k].cpu().numpy(), unit_scale=True, n_threads=k)
for sel in ("mma", "tc5"):
    o, v, _ = outs[sel]
    err = o[:k].astype(np.float64) - ref
    print(sel, "vs oracle: max", np.abs(err).max(), "snr", 10 * np.log10((ref.astype(np.float64) ** 2).mean() / (err ** 2).mean()), "vad", np.abs(v[:k] - rv).max())
