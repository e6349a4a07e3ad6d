"""CPU: host logic and the C-ABI boundary (no compute calls -- there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import crispy_b200 as cb
from crispy_b200 import _lib, build
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    build.build()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "crispy_ns.h")).read()
    declared = set(re.findall(r"\b(crispy_ns_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "header parse failed"
    L = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"libcrispy_ns.so lacks {missing}"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert L.crispy_ns_frame_size() == 480


def test_c_consumer_compiles_against_the_header_and_links(tmp_path):
    """tests/c_abi/abi_smoke.c is compiled as strict C99 against include/crispy_ns.h and linked to libcrispy_ns.so:
    a prototype drift between the header and the library (or the Python ctypes table) cannot hide.  Without a GPU
    the program checks the no-fallback contract and exits 0; the GPU suite runs it in full."""
    import subprocess
    exe = str(tmp_path / "abi_smoke")
    lib_dir = os.path.join(ROOT, "crispy_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi", "abi_smoke.c"), "-L" + lib_dir, "-lcrispy_ns", "-lm",
                           "-Wl,-rpath," + lib_dir, "-o", exe])
    if cb.device_count() == 0:
        r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
        assert r.returncode == 0 and "ENODEV contract ok" in r.stdout, r.stdout + r.stderr


def test_abi_wrappers_are_generated_from_the_header():
    """crispy_ns_abi.inc (the extern "C" try/catch forwarding layer) is what scripts/gen_abi.py makes of the header."""
    import subprocess
    import sys
    inc = os.path.join(ROOT, "crispy_b200", "csrc", "crispy_ns_abi.inc")
    before = open(inc).read()
    subprocess.check_call([sys.executable, os.path.join(ROOT, "scripts", "gen_abi.py")], stdout=subprocess.DEVNULL)
    assert open(inc).read() == before, "include/crispy_ns.h changed: run scripts/gen_abi.py and rebuild"
    assert before.count("guard(") + before.count("catch (...)") >= len(_lib.SYMBOLS)


def test_library_contains_sm_100a_code():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    # the Blackwell-only instructions are really in the binary: tcgen05 mma / commit / tensor-memory load and store in
    # the tcgen05 recurrent core, TMA bulk copies in the synthesis kernel, warp-level HMMA in the default recurrent core
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "HMMA.16816"):
        assert mnemonic in sass, mnemonic


def test_models_match_oracle_generator_and_loader():
    for seed in (0, 1, 12345):
        assert cb.Model.synthetic(seed).to_bytes() == po.Model.synthetic(seed).to_bytes()
    blob = cb.Model.synthetic(3).to_bytes()
    assert cb.Model.from_bytes(blob).to_bytes() == blob
    with pytest.raises(cb.CrispyNsError):
        cb.Model.from_bytes(b"not a model")
    with pytest.raises(cb.CrispyNsError):
        cb.Model.from_bytes(blob[:-5])


@pytest.mark.skipif(cb.device_count() > 0, reason="checks the no-device behaviour")
def test_no_cpu_fallback_without_a_device():
    with pytest.raises(cb.CrispyNsError, match="no CPU fallback"):
        cb.DenoiseState.new()
    with pytest.raises(cb.CrispyNsError, match="no CPU fallback"):
        cb.BatchDenoiser(4)


def test_python_linear_resampler_matches_oracle():  # audio.rs:73-134
    rng = np.random.default_rng(0)
    x = rng.standard_normal(2000).astype(np.float32)
    for rin, rout in ((44100.0, 48000.0), (48000.0, 16000.0), (16000.0, 48000.0), (48000.0, 48000.0)):
        r = cb.LinearResampler(rin, rout)
        got = []
        for s in x:
            r.process_sample(s, got.append)
        want = po.linear_resample(x, rin, rout)
        assert np.array_equal(np.array(got, np.float32), want)
        n = _lib.lib().crispy_ns_linear_resample_count(rin, rout, len(x))
        assert n == len(want)
    r = cb.LinearResampler(44100.0, 48000.0)  # audio.rs:1083-1096
    assert abs(r.rates()[0] - 44100.0) < 0.1 and abs(r.rates()[1] - 48000.0) < 0.1
    r.set_rates(44100.0, 16000.0)
    assert abs(r.rates()[1] - 16000.0) < 0.1


def test_wav_roundtrip_and_header(tmp_path):  # recording.rs:406-480
    p = str(tmp_path / "t.wav")
    left = np.full(100, (np.float32(0.5) * np.float32(32767.0)).astype(np.int16))
    right = np.full(100, -left[0])
    inter = np.stack([left, right], axis=1).astype(np.int16)
    cb.wav_write_pcm16(p, inter, channels=2, sample_rate=48000)
    raw = open(p, "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and len(raw) == 44 + 400
    assert int.from_bytes(raw[22:24], "little") == 2 and int.from_bytes(raw[24:28], "little") == 48000
    assert int.from_bytes(raw[34:36], "little") == 16
    data, sr = cb.wav_read_pcm16(p)
    assert sr == 48000 and np.array_equal(data, inter)
    with pytest.raises(cb.CrispyNsError, match="length mismatch"):  # recording.rs:102-104
        cb.wav_write_pcm16(p, np.zeros(3, np.int16), channels=2)
    with pytest.raises(cb.CrispyNsError):
        cb.wav_read_pcm16(str(tmp_path / "missing.wav"))


def test_wav_reader_walks_chunks_like_get_wav_duration(tmp_path):  # commands/recording.rs:385-460
    """The files either side of the path: what hound 3.5.1 writes for the recorder's spec (recording.rs:85-90: the
    plain 44-byte PCM header), what ffmpeg's import writes (commands/convert.rs:113-127: a LIST/INFO chunk between
    fmt and data), an 18-byte fmt chunk, an odd-sized chunk with its pad byte -- all read back sample for sample;
    anything that is not 16-bit PCM, a truncated data chunk and a non-RIFF file are refused with CRISPY_NS_EIO; a
    recording that never reached finalize() (hound leaves data size 0) has no frames, where the reference's walk
    returns None."""
    import struct
    q = (np.arange(200, dtype=np.int16) - 100).reshape(100, 2)

    def fmt(n=16, tag=1, ch=2, sr=48000, bits=16, extra=b""):
        return b"fmt " + struct.pack("<IHHIIHH", n, tag, ch, sr, sr * ch * bits // 8, ch * bits // 8, bits) + extra

    def wav(chunks, data):
        body = b"WAVE" + b"".join(chunks) + b"data" + struct.pack("<I", len(data)) + data
        return b"RIFF" + struct.pack("<I", len(body)) + body

    lavf = b"LIST" + struct.pack("<I", 26) + b"INFOISFT" + struct.pack("<I", 14) + b"Lavf60.16.100\0"
    guid = lambda t: struct.pack("<H", t) + bytes.fromhex("000000001000800000aa00389b71")  # KSDATAFORMAT_SUBTYPE_*
    ext = struct.pack("<HHI", 22, 16, 3)  # cbSize, valid bits, channel mask
    streamed = wav([fmt()], q.tobytes())
    streamed = streamed[:40] + struct.pack("<I", 0xFFFFFFFF) + streamed[44:]  # ffmpeg to a pipe: size never patched
    good = {"hound44": wav([fmt()], q.tobytes()),
            "extensible_pcm": wav([fmt(40, tag=0xFFFE, extra=ext + guid(1))], q.tobytes()),
            "streamed_size": streamed,
            "fmt17_padded": wav([fmt(17, extra=b"\0\0")], q.tobytes()),
            "ffmpeg_list": wav([fmt(), lavf], q.tobytes()),
            "fmt18": wav([fmt(18, extra=b"\0\0")], q.tobytes()),
            "odd_chunk_padded": wav([fmt(), b"junk" + struct.pack("<I", 3) + b"abc\0"], q.tobytes())}
    assert len(good["hound44"]) == 44 + 400
    for name, blob in good.items():
        p = tmp_path / (name + ".wav")
        p.write_bytes(blob)
        y, sr = cb.wav_read_pcm16(str(p))
        assert sr == 48000 and np.array_equal(y, q), name
    p = tmp_path / "never_finalized.wav"
    p.write_bytes(wav([fmt()], b""))
    y, sr = cb.wav_read_pcm16(str(p))
    assert y.shape == (0, 2) and sr == 48000
    bad = {"float32": wav([fmt(tag=3, bits=32)], np.zeros(200, np.float32).tobytes()),
           "extensible_float": wav([fmt(40, tag=0xFFFE, bits=32, extra=struct.pack("<HHI", 22, 32, 3) + guid(3))],
                                   np.zeros(200, np.float32).tobytes()),
           "pcm24": wav([fmt(bits=24)], bytes(600)),
           "truncated": wav([fmt()], q.tobytes())[:-50],
           "not_riff": b"RIFX" + bytes(60),
           "no_data_chunk": wav([fmt()], q.tobytes())[:36]}
    for name, blob in bad.items():
        p = tmp_path / (name + ".wav")
        p.write_bytes(blob)
        with pytest.raises(cb.CrispyNsError):
            cb.wav_read_pcm16(str(p))


def test_sinc_front_end_geometry_and_argument_checks():  # no device needed: validation comes first
    L = _lib.lib()
    assert L.crispy_ns_sinc_resample_count(44100, 48000, 441) == 480
    assert L.crispy_ns_sinc_resample_count(48000, 16000, 300) == 100
    assert L.crispy_ns_sinc_resample_count(44100, 48000, 0) == 0
    assert L.crispy_ns_sinc_resample_count(44101, 48000, 100) == 0  # L = 48000 > 1024: unsupported ratio
    assert cb.sinc_needed(44100, 48000, 441000, 4800, 4800) == (4283, 4665)
    assert cb.sinc_needed(44100, 48000, 441000, 0, 480) == (0, 441 + 128)  # clipped at the start of the recording
    assert cb.sinc_needed(44100, 48000, 441, 0, 480) == (0, 441)            # and at its end
    fake = 16  # never dereferenced: the argument checks fail first
    assert L.crispy_ns_sinc_resample_chunk(0, fake, 0, 441, 441, fake, 7, 160, 1, 441, 480, 44100, 48000, 0, 0.0, None) == -1
    assert b"multiple of L" in L.crispy_ns_last_error()
    assert L.crispy_ns_sinc_resample_chunk(0, fake, 0, 441, 441, fake, 0, 481, 1, 441, 481, 44100, 48000, 0, 0.0, None) == -1
    assert L.crispy_ns_sinc_resample_chunk(0, fake, 100, 341, 441, fake, 0, 480, 1, 441, 480, 44100, 48000, 0, 0.0, None) == -1
    assert b"window" in L.crispy_ns_last_error()
    assert L.crispy_ns_sinc_resample(0, fake, fake, 1, 441, 441, 480, 44100, 48000, 7, 0.0, None) == -1  # odd sinc_len



def test_resample_audio_count_matches_the_oracle():  # recording.rs:19, :27-35; host arithmetic only
    import numpy as np
    from oracle import pyoracle as po
    L = _lib.lib()
    for n, fr, to in ((0, 44100, 48000), (1, 44100, 48000), (441, 44100, 48000), (44101, 44100, 48000), (4800, 48000, 44100),
                      (999, 48000, 16000), (1000, 16000, 48000), (777, 48000, 48000), (12345, 22050, 48000), (7, 8000, 48000)):
        assert L.crispy_ns_resample_audio_count(n, fr, to) == len(po.resample_audio(np.zeros(n, np.float32), fr, to)), (n, fr, to)
    assert L.crispy_ns_resample_audio_count(100, 0, 48000) == 0


def test_rust_shim_extern_block_matches_the_header():
    """bindings/rust/ns_gpu.rs cannot be compiled here (no rustc), so its `extern "C"` block is checked against
    include/crispy_ns.h token by token: every function it binds is declared, with the same number of arguments and
    the same type for each (int -> c_int, float -> c_float, int64_t -> i64, uint32_t -> u32, size_t -> usize,
    pointers with their const-ness), and the flag constants carry the header's values."""
    import re
    h = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "crispy_ns.h")).read(), flags=re.S)
    rs = open(os.path.join(ROOT, "bindings", "rust", "ns_gpu.rs")).read()
    scalar = {"int": "c_int", "float": "c_float", "int64_t": "i64", "uint32_t": "u32", "size_t": "usize",
              "uint64_t": "u64", "double": "f64", "void": "c_void", "char": "c_char", "int16_t": "i16"}
    opaque = {"crispy_ns_model": "CrispyNsModel", "crispy_ns_state": "CrispyNsState",
              "crispy_ns_batch": "CrispyNsBatch", "crispy_ns_multi": "CrispyNsMulti"}

    def c_to_rust(decl: str) -> str:
        decl = re.sub(r"\b[A-Za-z_]\w*\s*$", "", decl.strip()) if not decl.strip().endswith("*") else decl.strip()
        toks = re.findall(r"const|\*|\w+", decl)
        base = [t for t in toks if t not in ("const", "*", "struct")][0]
        out = opaque.get(base) or scalar[base]
        # walk the declarator left to right: "const T *" -> *const T, "T *" -> *mut T, "const T *const *" -> *const *const T
        consts, stars, pending = [], 0, "const" in toks[:toks.index(base) + 1]
        for t in toks[toks.index(base) + 1:]:
            if t == "*":
                consts.append(pending)
                pending = False
            elif t == "const":
                pending = True
        for is_const in consts:
            out = ("*const " if is_const else "*mut ") + out
        return out

    protos = {}
    for m in re.finditer(r"([\w \*]+?)\b(crispy_ns_\w+)\s*\(([^;{]*?)\)\s*;", h, flags=re.S):
        args = " ".join(m.group(3).split())
        protos[m.group(2)] = ([] if args in ("", "void") else [c_to_rust(a) for a in args.split(",")],
                              " ".join(m.group(1).split()))
    blk = rs[rs.index('extern "C" {'):]
    blk = blk[:blk.index("\n}\n")]
    bound = 0
    for m in re.finditer(r"fn (crispy_ns_\w+)\s*\((.*?)\)\s*(?:->\s*([\w\* ]+?))?\s*;", blk, flags=re.S):
        name, args, ret = m.group(1), m.group(2).strip().rstrip(","), (m.group(3) or "").strip()
        assert name in protos, f"{name} is not declared in crispy_ns.h"
        rust_types = [" ".join(a.split(":", 1)[1].split()) for a in args.split(",")] if args else []
        assert rust_types == protos[name][0], (name, rust_types, protos[name][0])
        c_ret = protos[name][1]
        want_ret = "" if c_ret == "void" else c_to_rust(c_ret + " x" if not c_ret.endswith("*") else c_ret)
        assert ret == want_ret, (name, ret, want_ret)
        bound += 1
    assert bound >= 20
    for flag in ("IN_I16", "OUT_I16", "UNIT_SCALE", "MIX_STEREO_I16", "DROP_FIRST_FRAME"):
        hv = re.search(rf"CRISPY_NS_{flag}\s*=\s*1u\s*<<\s*(\d+)", h).group(1)
        rv = re.search(rf"CRISPY_NS_{flag}: u32 = 1 << (\d+)", rs).group(1)
        assert hv == rv, flag
    # the drop-in surface audio.rs type-checks against (audio.rs:4, :203, :229, :268)
    assert "pub const FRAME_SIZE: usize = 480;" in rs
    assert "pub struct DenoiseState<'model>" in rs and "pub fn new() -> Box<DenoiseState<'static>>" in rs
    assert "pub fn process_frame(&mut self, output: &mut [f32], input: &[f32]) -> f32" in rs


def test_ctypes_prototypes_match_the_header():
    """crispy_b200/_lib.py declares argtypes by hand: every prototype in include/crispy_ns.h is compared with it,
    argument by argument (scalar width and signedness, pointer or not) and for the return type, so a drift between
    the header and the Python binding cannot go unseen."""
    import ctypes as C
    import re
    h = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "crispy_ns.h")).read(), flags=re.S)
    L = _lib.lib()
    scalar = {"int": C.c_int, "float": C.c_float, "int64_t": C.c_int64, "uint32_t": C.c_uint32, "size_t": C.c_size_t,
              "uint64_t": C.c_uint64, "double": C.c_double}

    def is_pointer(t) -> bool:
        return t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer))

    seen = 0
    for m in re.finditer(r"([\w \*]+?)\b(crispy_ns_\w+)\s*\(([^;{]*?)\)\s*;", h, flags=re.S):
        ret, name, args = " ".join(m.group(1).split()), m.group(2), " ".join(m.group(3).split())
        if "typedef" in ret:
            continue
        fn = getattr(L, name)
        c_args = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        if c_args:
            assert fn.argtypes is not None and len(fn.argtypes) == len(c_args), (name, fn.argtypes, c_args)
        for a, t in zip(c_args, fn.argtypes or []):
            if "*" in a:
                assert is_pointer(t), (name, a, t)
            else:
                base = [w for w in re.findall(r"\w+", a) if w != "const"][0]
                assert t is scalar[base], (name, a, t)
        if ret == "void":
            assert fn.restype is None, name
        elif "*" in ret:
            assert is_pointer(fn.restype), (name, fn.restype)
        else:
            assert fn.restype is scalar[ret], (name, ret, fn.restype)
        seen += 1
    assert seen == len(_lib.SYMBOLS), (seen, len(_lib.SYMBOLS))


def test_product_never_touches_the_oracle_and_has_no_cpu_fallback():
    """The oracle is test infrastructure: nothing under crispy_b200/ (Python or C++/CUDA), include/ or bindings/ may
    import, include, link or load it; importing the product must not pull the oracle's modules in either; and the
    library links neither the oracle nor the host emulation (its only compute is the sm_100a code in it)."""
    import re
    import subprocess
    import sys
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b|#\s*include\s*[<\"][^>\"]*(oracle|ns_emu)[^>\"]*[>\"])|"
                     r"(CDLL|dlopen)\([^)]*oracle", re.M)
    for top in ("crispy_b200", "include", "bindings"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".inc", ".rs")):
                    src = open(os.path.join(dirpath, f), errors="replace").read()
                    assert not pat.search(src), os.path.join(dirpath, f)
    code = ("import sys; sys.path.insert(0, %r); import crispy_b200, crispy_b200.denoise, crispy_b200.shard; "
            "bad = [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]; print(bad); sys.exit(bool(bad))" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "ns_emu" not in ldd, ldd
