"""The sinc front end alone (for ncu): 256 streams x 10 s at 44.1 kHz -> 48 kHz, a few launches."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import crispy_b200 as cb
x = torch.randn((256, 441000), device="cuda")
for _ in range(3):
    y = cb.sinc_resample(x, 44100, 48000)
torch.cuda.synchronize()
print(y.shape)
