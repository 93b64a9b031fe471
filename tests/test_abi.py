"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol the header declares.
No compute call is made (there is no GPU here)."""
import os
import re

import pytest

from vae_tagger_b200 import _build, _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _build.build()
    return _native.load_library()


def header_functions():
    src = open(os.path.join(ROOT, "include", "vae_tagger_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib):
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vae_tagger_b200.h but not exported"
        assert n in _native.SYMBOLS, f"{n} has no ctypes signature"
    assert sorted(_native.SYMBOLS) == names


def test_abi_version_and_error_string(lib):
    assert lib.vt_abi_version() == 1
    assert isinstance(lib.vt_last_error(), bytes)


def test_struct_sizes_match_header_layout():
    import ctypes as C

    import shutil
    import subprocess

    got = [C.sizeof(t) for t in (_native.EncoderConfig, _native.HeadConfig, _native.EncodeArgs, _native.TagArgs,
                                 _native.InferHostArgs, _native.HeadTrainArgs, _native.ResizeArgs, _native.DecodeArgs, _native.InferArgs)]
    assert got == [72, 32, 96, 72, 80, 136, 80, 56, 80]
    gcc = shutil.which("gcc")
    if gcc:  # compile the header as C and compare sizeof() of every struct
        import tempfile

        with tempfile.TemporaryDirectory() as d:
            src = os.path.join(d, "sz.c")
            open(src, "w").write(
                '#include <stdio.h>\n#include "vae_tagger_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu",'
                "sizeof(vt_encoder_config),sizeof(vt_head_config),sizeof(vt_encode_args),sizeof(vt_tag_args),"
                "sizeof(vt_infer_host_args),sizeof(vt_head_train_args),sizeof(vt_resize_args),sizeof(vt_decode_args),sizeof(vt_infer_args));return 0;}\n")
            exe = os.path.join(d, "sz")
            subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
            out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
        assert [int(v) for v in out] == got


def test_sass_contains_blackwell_instructions():
    """The library must contain tcgen05 MMA / TMEM loads / TMA (SASS mnemonics UTC*MMA, LDTM, UTMALDG)."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    _build.build()
    sass = subprocess.run([cuobjdump, "-sass", _build.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass
    assert "LDTM" in sass
    assert "UTMALDG" in sass
