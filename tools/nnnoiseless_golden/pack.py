#!/usr/bin/env python
"""c1_input.f32 + c1_output.f32 + c1_vad.f32 (written by the Rust generator) -> tests/golden/nnnoiseless_c1.npz,
the fixture tests/test_real_parity.py looks for:  python tools/nnnoiseless_golden/pack.py [dir_with_the_three_files]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    d = sys.argv[1] if len(sys.argv) > 1 else "."
    x = np.fromfile(os.path.join(d, "c1_input.f32"), "<f4")
    out = np.fromfile(os.path.join(d, "c1_output.f32"), "<f4")
    vad = np.fromfile(os.path.join(d, "c1_vad.f32"), "<f4")
    assert x.size == out.size and x.size == vad.size * 480, (x.size, out.size, vad.size)
    dst = os.path.join(ROOT, "tests", "golden", "nnnoiseless_c1.npz")
    np.savez_compressed(dst, x=x, out=out, vad=vad, producer=np.array("nnnoiseless 0.5.2 DenoiseState::new + process_frame"))
    print(f"wrote {dst}: {vad.size} frames")
