"""Parity with the REAL reference, nnnoiseless 0.5.2 -- runs by itself as soon as its two inputs exist:

  tests/golden/nnnoiseless_c1.npz    the seeded configs[0] clip and what the crate's process_frame returned for it
  weights/rnnoise_default.crnsmdl    the crate's built-in weights
(both made by tools/nnnoiseless_golden/ on a machine with cargo + network; $CRISPY_NNNOISELESS_GOLDEN and
$CRISPY_NS_REAL_WEIGHTS override the paths).  Until then the two real tests SKIP, loudly: nothing in this repository
has been compared with nnnoiseless itself (parity unpinned).  The remaining tests keep the
machinery honest without the files: the comparison code runs on a stand-in fixture made by the oracle, and the
rnn_data.c converter round-trips the synthetic model.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "nnnoiseless_golden"))

from oracle import pyoracle as po  # noqa: E402
from tests.util import TOL_MAX_ABS, TOL_SNR_DB, TOL_VAD, make_signal, parity_report  # noqa: E402

GOLDEN = os.environ.get("CRISPY_NNNOISELESS_GOLDEN", os.path.join(ROOT, "tests", "golden", "nnnoiseless_c1.npz"))
WEIGHTS = os.environ.get("CRISPY_NS_REAL_WEIGHTS", os.path.join(ROOT, "weights", "rnnoise_default.crnsmdl"))
HAVE = os.path.exists(GOLDEN) and os.path.exists(WEIGHTS)
WHY = ("PARITY UNPINNED: no nnnoiseless 0.5.2 output / weights in the tree (" + os.path.relpath(GOLDEN, ROOT) + ", "
       + os.path.relpath(WEIGHTS, ROOT) + "); make them with tools/nnnoiseless_golden/README.md")


def check_against(golden_path: str, run) -> dict:
    """run(x [1, n] f32 in 16-bit scale) -> (out [1, n], vad [1, frames]); north_star tolerances against the fixture"""
    g = np.load(golden_path)
    x = g["x"].astype(np.float32)[None, :]
    out, vad = run(x)
    r = parity_report(g["out"][None, :], out, g["vad"][None, :], vad)
    assert r["max_abs"] <= TOL_MAX_ABS, r
    assert r["snr_db"] >= TOL_SNR_DB, r
    assert r["vad_max"] <= TOL_VAD, r
    return r


@pytest.mark.skipif(not HAVE, reason=WHY)
def test_oracle_matches_nnnoiseless():
    """The oracle, loaded with the crate's weights, against the crate's own output: this is what pins the oracle."""
    model = po.Model.from_bytes(open(WEIGHTS, "rb").read())
    print("oracle vs nnnoiseless 0.5.2:", check_against(GOLDEN, lambda x: po.process_streams(model, x)))


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE, reason=WHY)
def test_cuda_path_matches_nnnoiseless():
    """The CUDA path through the C ABI, loaded with the crate's weights, against the crate's own output."""
    import torch
    import crispy_b200 as cb
    model = cb.Model.from_bytes(open(WEIGHTS, "rb").read())

    def run(x):
        den = cb.BatchDenoiser(1, model)
        out, vad = den.process_streams(torch.from_numpy(x).cuda(), unit_scale=False)
        return out.cpu().numpy(), vad.cpu().numpy()

    print("CUDA vs nnnoiseless 0.5.2:", check_against(GOLDEN, run))


def test_real_parity_machinery_on_a_stand_in_fixture(tmp_path, oracle_model):
    """No claim about nnnoiseless: a fixture of the same shape made by the ORACLE exercises the comparison code, and
    a perturbed copy must fail it."""
    x = make_signal(2, 120)[1]
    out, vad = po.process_streams(oracle_model, x[None, :])
    p = str(tmp_path / "stand_in.npz")
    np.savez_compressed(p, x=x, out=out[0], vad=vad[0])
    r = check_against(p, lambda xx: po.process_streams(oracle_model, xx))
    assert r["max_abs"] == 0.0
    bad = out[0].copy()
    bad[5000] += 40.0  # > 1e-3 of full scale
    np.savez_compressed(p, x=x, out=bad, vad=vad[0])
    with pytest.raises(AssertionError):
        check_against(p, lambda xx: po.process_streams(oracle_model, xx))


def test_rnn_data_converter_round_trips_a_model(oracle_model):
    """tools/nnnoiseless_golden/rnn_data_to_crnsmdl.py on a rnn_data.c-shaped text of the synthetic model gives back
    the model's CRNSMDL1 bytes, which both loaders (oracle and library) accept."""
    import struct

    import rnn_data_to_crnsmdl as conv
    blob = oracle_model.to_bytes()
    off, text = 8, ["/* shaped like xiph/rnnoise src/rnn_data.c */", '#include "rnn.h"']
    names = {0: "ACTIVATION_TANH", 1: "ACTIVATION_SIGMOID", 2: "ACTIVATION_RELU"}
    for name, kind, n_in, n in conv.LAYERS:
        k, i, o, act = struct.unpack_from("<IIII", blob, off)
        assert (k, i, o) == (kind, n_in, n)
        off += 16
        sizes = ([("weights", n_in * n), ("bias", n)] if kind == 0 else
                 [("weights", n_in * 3 * n), ("recurrent_weights", n * 3 * n), ("bias", 3 * n)])
        for suffix, cnt in sizes:
            vals = struct.unpack_from(f"{cnt}b", blob, off)
            off += cnt
            text.append(f"static const rnn_weight {name}_{suffix}[{cnt}] = {{\n   " + ", ".join(map(str, vals)) + "};")
        if kind == 0:
            text.append(f"const DenseLayer {name} = {{\n   {name}_bias,\n   {name}_weights,\n   {n_in}, {n}, {names[act]}\n}};")
        else:
            text.append(f"const GRULayer {name} = {{\n   {name}_bias,\n   {name}_weights,\n   {name}_recurrent_weights,\n   {n_in}, {n}, {names[act]}\n}};")
    assert off == len(blob)
    assert conv.convert("\n\n".join(text)) == blob
    import crispy_b200 as cb
    assert cb.Model.from_bytes(blob).to_bytes() == blob
