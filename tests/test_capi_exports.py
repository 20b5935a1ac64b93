"""The C-ABI library loads and exports every symbol include/heic_b200.h declares (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "heic_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(heic_b200_\w+)\s*\(", text)))


def test_header_symbols_are_exported(built):
    from heif_b200 import _capi

    lib = C.CDLL(_capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/heic_b200.h but not exported"
    bound = {n for n, _, _ in _capi.SYMBOLS}
    assert bound == set(names), f"ctypes table and header disagree: {bound ^ set(names)}"
    assert lib.heic_b200_abi_version() == 1


def test_struct_sizes_match_the_header(built):
    """ctypes mirrors vs a tiny C program compiled against the header."""
    import subprocess
    import tempfile

    from heif_b200 import _capi as K

    src = r'''
#include <stdio.h>
#include "heic_b200.h"
int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(heic_sps), sizeof(heic_pps), sizeof(heic_slice_header),
  sizeof(heic_tile_desc), sizeof(heic_image_desc), sizeof(heic_tile_status), sizeof(heic_file_info), sizeof(heic_tile_dump)); return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "include"), "-o", os.path.join(d, "s"), os.path.join(d, "s.c")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "s")]).split()]
    mine = [C.sizeof(t) for t in (K.Sps, K.Pps, K.SliceHeader, K.TileDesc, K.ImageDesc, K.TileStatus, K.FileInfo, K.TileDump)]
    assert mine == sizes


def test_no_device_is_an_error_code_not_a_fallback(built):
    """Without a CUDA device the compute entry points fail loudly (HEIC_E_NO_DEVICE); there is no CPU path."""
    import heif_b200 as H

    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(H.HeicError) as e:
        H.HeicDecoder()
    assert e.value.code == H._capi.HEIC_E_NO_DEVICE


def test_rust_sys_crate_declares_every_header_entry_point():
    """rust/heic-b200-sys is source only (no Rust toolchain here), so at least its extern block must name exactly the
    functions include/heic_b200.h declares."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "heic_b200.h")).read()
    rust = open(os.path.join(root, "rust", "heic-b200-sys", "src", "lib.rs")).read()
    in_header = set(re.findall(r"\b(heic_b200_[a-z_0-9]+)\s*\(", header))
    in_rust = set(re.findall(r"\bfn (heic_b200_[a-z_0-9]+)", rust))
    assert in_header == in_rust, (sorted(in_header - in_rust), sorted(in_rust - in_header))
