"""The C-ABI shared library loads without a GPU and exports every symbol include/vqw.h
declares (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vqw.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vqw_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "chainer_vq_vae_b200", "csrc", "libvqw.so"))
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vqw.h but not exported"
    lib.vqw_version.restype = ctypes.c_int
    assert lib.vqw_version() >= 100
    lib.vqw_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.vqw_last_error(), bytes)


def test_python_binding_covers_the_header():
    import chainer_vq_vae_b200 as V
    bound = set(V._lib._SIGNATURES)
    assert set(_declared()) <= bound, set(_declared()) - bound


def test_argument_errors_are_reported_before_launch():
    """Negative return + message, no CUDA call needed (runs on the CPU box)."""
    import chainer_vq_vae_b200 as V
    lib = V._lib.lib
    rc = lib.vqw_vq_forward(None, None, None, None, None, None, None, 1, 0, 4, 8, None)
    assert rc < 0 and b"bad sizes" in lib.vqw_last_error()
    d = V._lib.ResblockDesc()
    d.B, d.T, d.Cr, d.Cd, d.Cs, d.Cc, d.fs, d.dilation = 1, 8, 32, 33, 32, 16, 3, 1
    w = V._lib.ResblockWeights()
    rc = lib.vqw_resblock_forward(ctypes.byref(d), None, None, ctypes.byref(w), None, None, None,
                                  None, None)
    assert rc < 0 and b"even" in lib.vqw_last_error()
    assert lib.vqw_resnet_forward_workspace(None) == -1


def test_no_oracle_or_cpu_fallback_in_the_product():
    """The product package must not import the oracle or route around the CUDA library."""
    pkg = os.path.join(ROOT, "chainer_vq_vae_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), fn


def test_arithmetic_modes_agree_between_header_binding_and_bench():
    """include/vqw.h's VQW_MODE_* constants, _lib.MODES and bench.py's --mode choices name the same
    set of arithmetic modes with the same values."""
    import re
    import chainer_vq_vae_b200._lib as L
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "vqw.h")).read()
    in_header = {m.group(1).lower(): int(m.group(2))
                 for m in re.finditer(r"#define\s+VQW_MODE_(\w+)\s+(\d+)", hdr)}
    assert in_header == dict(L.MODES), (in_header, L.MODES)
    bench = open(os.path.join(root, "bench.py")).read()
    choices = re.search(r'"--mode".*?choices=\[(.*?)\]', bench, re.S).group(1)
    assert {c.strip().strip('"') for c in choices.split(",")} == set(L.MODES)
    assert set(L.X3_MODES) == {L.MODES["bf16x3"], L.MODES["fp16x3"]}
