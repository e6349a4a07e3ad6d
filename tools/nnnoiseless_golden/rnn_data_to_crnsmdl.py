#!/usr/bin/env python
"""xiph/rnnoise src/rnn_data.c (the 2018 default weights nnnoiseless embeds) -> weights/rnnoise_default.crnsmdl.

nnnoiseless keeps its weights private to the crate, so the blob is made from the public C file they were generated
from (https://github.com/xiph/rnnoise, BSD; tag v0.1 / commit of 2018, `src/rnn_data.c`):

    python tools/nnnoiseless_golden/rnn_data_to_crnsmdl.py path/to/rnn_data.c [out.crnsmdl]

CRNSMDL1 layout (crispy_b200/csrc/ns_host.cpp model_from_bytes): magic "CRNSMDL1", then input_dense, vad_gru,
vad_output, noise_gru, denoise_gru, denoise_output; each layer = u32 kind (0 dense, 1 GRU), u32 inputs, u32 neurons,
u32 activation (0 tanh, 1 sigmoid, 2 ReLU), then int8 arrays: dense weights[in*out], bias[out]; GRU
input_weights[in*3n], recurrent_weights[n*3n], bias[3n] -- the same element order as the C arrays.
"""
import os
import re
import struct
import sys

LAYERS = [  # name, kind, inputs, neurons
    ("input_dense", 0, 42, 24), ("vad_gru", 1, 24, 24), ("vad_output", 0, 24, 1),
    ("noise_gru", 1, 90, 48), ("denoise_gru", 1, 114, 96), ("denoise_output", 0, 96, 22),
]
ACT = {"ACTIVATION_TANH": 0, "ACTIVATION_SIGMOID": 1, "ACTIVATION_RELU": 2}


def parse(src: str):
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    arrays = {}
    for m in re.finditer(r"(\w+)\s*\[\s*\d*\s*\]\s*=\s*\{([^}]*)\}", src):
        body = m.group(2).strip()
        if re.fullmatch(r"[-\d,\s]*", body):
            arrays[m.group(1)] = [int(v) for v in re.findall(r"-?\d+", body)]
    acts = {}
    for m in re.finditer(r"(?:DenseLayer|GRULayer)\s+(\w+)\s*=\s*\{([^}]*)\}", src):
        a = re.findall(r"ACTIVATION_\w+", m.group(2))
        if a:
            acts[m.group(1)] = ACT[a[-1]]
    return arrays, acts


def convert(src: str) -> bytes:
    arrays, acts = parse(src)
    out = bytearray(b"CRNSMDL1")

    def arr(name, n):
        a = arrays.get(name)
        if a is None or len(a) != n:
            raise SystemExit(f"rnn_data.c: array {name} missing or of length {None if a is None else len(a)} != {n}")
        if any(v < -128 or v > 127 for v in a):
            raise SystemExit(f"rnn_data.c: {name} holds values outside int8")
        return struct.pack(f"{n}b", *a)

    for name, kind, n_in, n in LAYERS:
        if name not in acts:
            raise SystemExit(f"rnn_data.c: layer initialiser {name} not found")
        out += struct.pack("<IIII", kind, n_in, n, acts[name])
        if kind == 0:
            out += arr(name + "_weights", n_in * n) + arr(name + "_bias", n)
        else:
            out += arr(name + "_weights", n_in * 3 * n) + arr(name + "_recurrent_weights", n * 3 * n) + arr(name + "_bias", 3 * n)
    return bytes(out)


if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(root, "weights", "rnnoise_default.crnsmdl")
    blob = convert(open(sys.argv[1]).read())
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    open(dst, "wb").write(blob)
    print(f"wrote {dst}: {len(blob)} bytes ({len(blob) - 8 - 16 * 6} int8 parameters)")
