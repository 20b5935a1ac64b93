"""ctypes wrapper around oracle/_build/liboracle.so (hevc_oracle.c).  TEST INFRASTRUCTURE ONLY.

Built by oracle/Makefile (`make -C oracle`) or __graft_entry__.build().
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from heif_b200 import _capi as K

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


class OracleOut(C.Structure):
    _fields_ = [
        ("plane", C.c_void_p * 3), ("recon", C.c_void_p * 3), ("deblocked", C.c_void_p * 3),
        ("tu_map", C.c_void_p), ("level", C.c_void_p * 3), ("resid", C.c_void_p * 3),
        ("qp_map", C.c_void_p), ("sao", C.c_void_p),
        ("bins", C.c_uint32), ("ctus", C.c_uint32), ("error", C.c_char * 160),
    ]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        args = [C.POINTER(K.Sps), C.POINTER(K.Pps), C.POINTER(K.SliceHeader), C.c_void_p, C.c_uint32, C.POINTER(OracleOut)]
        lib.hevc_oracle_decode_picture.argtypes = args
        lib.hevc_oracle_parse_picture.argtypes = args
        lib.hevc_oracle_color_stitch.argtypes = [C.c_void_p] + [C.c_uint32] * 8 + [C.c_void_p, C.c_uint64]
        lib.hevc_oracle_color_stitch.restype = None
        lib.hevc_oracle_idct.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.hevc_oracle_idct.restype = None
        lib.hevc_oracle_context_init.argtypes = [C.c_int, C.c_void_p]
        lib.hevc_oracle_context_init.restype = None
        lib.hevc_oracle_test_binarization.argtypes = [C.c_int, C.c_uint32, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        lib.hevc_oracle_test_binarization.restype = C.c_uint32
        lib.hevc_oracle_tu_map_len.argtypes = [C.POINTER(K.Sps)]
        lib.hevc_oracle_tu_map_len.restype = C.c_uint32
        lib.hevc_oracle_coeff_len.argtypes = [C.POINTER(K.Sps), C.c_int]
        lib.hevc_oracle_coeff_len.restype = C.c_uint32
        _lib = lib
    return _lib


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


def decode_picture(sps, pps, sh, rbsp, parse_only=False, intermediates=True):
    """Decode one picture.  rbsp: bytes / ctypes pointer + length tuple.  Returns dict of numpy arrays."""
    lib = load()
    if isinstance(rbsp, tuple):
        ptr, n = rbsp
        buf = C.cast(ptr, C.c_void_p)
    else:
        n = len(rbsp)
        keep = C.create_string_buffer(bytes(rbsp), n)
        buf = C.cast(keep, C.c_void_p)
    w, h = sps.pic_width_in_luma_samples, sps.pic_height_in_luma_samples
    chroma = sps.chroma_format_idc == 1
    out = OracleOut()
    res = {}

    def plane_set(name):
        arrs = [np.zeros((h, w), np.uint8)]
        if chroma:
            arrs += [np.zeros((h // 2, w // 2), np.uint8), np.zeros((h // 2, w // 2), np.uint8)]
        for i, a in enumerate(arrs):
            getattr(out, name)[i] = a.ctypes.data
        res[name] = arrs

    plane_set("plane")
    if intermediates:
        plane_set("recon")
        plane_set("deblocked")
        n_tu = lib.hevc_oracle_tu_map_len(C.byref(sps))
        res["tu_map"] = np.zeros(n_tu, np.uint32)
        out.tu_map = res["tu_map"].ctypes.data
        res["level"], res["resid"] = [], []
        for c in range(3 if chroma else 1):
            ln = lib.hevc_oracle_coeff_len(C.byref(sps), c)
            for name in ("level", "resid"):
                a = np.zeros(ln, np.int16)
                getattr(out, name)[c] = a.ctypes.data
                res[name].append(a)
        res["qp_map"] = np.zeros((h // 8, w // 8), np.uint8)
        out.qp_map = res["qp_map"].ctypes.data
        log2_ctb = sps.log2_min_luma_coding_block_size_minus3 + 3 + sps.log2_diff_max_min_luma_coding_block_size
        n_ctb = -(-w >> log2_ctb) * -(-h >> log2_ctb)
        res["sao"] = np.zeros(n_ctb * 4, np.uint32)
        out.sao = res["sao"].ctypes.data
    fn = lib.hevc_oracle_parse_picture if parse_only else lib.hevc_oracle_decode_picture
    rc = fn(C.byref(sps), C.byref(pps), C.byref(sh), buf, n, C.byref(out))
    if rc < 0:
        raise OracleError(rc, out.error.decode("utf-8", "replace"))
    res["bins"] = out.bins
    res["ctus"] = out.ctus
    return res


def decode_picture_into(sps, pps, sh, rbsp_ptr, rbsp_len, planes: np.ndarray):
    """Final planes only, written straight into `planes` (uint8, Y then Cb then Cr, w*h*3/2 bytes): nothing but the C call
    happens here, which is what bench.py's CPU arm times."""
    lib = load()
    w, h = sps.pic_width_in_luma_samples, sps.pic_height_in_luma_samples
    out = OracleOut()
    base = planes.ctypes.data
    out.plane[0] = base
    if sps.chroma_format_idc == 1:
        out.plane[1] = base + w * h
        out.plane[2] = base + w * h + (w // 2) * (h // 2)
    rc = lib.hevc_oracle_decode_picture(C.byref(sps), C.byref(pps), C.byref(sh), C.cast(rbsp_ptr, C.c_void_p), rbsp_len, C.byref(out))
    if rc < 0:
        raise OracleError(rc, out.error.decode("utf-8", "replace"))
    return out.bins


def color_stitch_into(planes_ptr: int, grid_rows, grid_cols, tile_w, tile_h, out_w, out_h, full_range, matrix_coeffs, rgb_ptr: int,
                      pitch: int):
    load().hevc_oracle_color_stitch(planes_ptr, grid_rows, grid_cols, tile_w, tile_h, out_w, out_h, full_range, matrix_coeffs,
                                    rgb_ptr, pitch)


def color_stitch(planes: np.ndarray, grid_rows, grid_cols, tile_w, tile_h, out_w, out_h, full_range=1, matrix_coeffs=6):
    lib = load()
    planes = np.ascontiguousarray(planes, dtype=np.uint8)
    rgb = np.zeros((out_h, out_w, 3), np.uint8)
    lib.hevc_oracle_color_stitch(planes.ctypes.data, grid_rows, grid_cols, tile_w, tile_h, out_w, out_h,
                                 full_range, matrix_coeffs, rgb.ctypes.data, out_w * 3)
    return rgb


def idct(coeff: np.ndarray, log2_size: int, dst: bool = False) -> np.ndarray:
    lib = load()
    coeff = np.ascontiguousarray(coeff, dtype=np.int16)
    out = np.zeros_like(coeff)
    lib.hevc_oracle_idct(coeff.ctypes.data, out.ctypes.data, log2_size, int(dst))
    return out


def context_init(slice_qp: int) -> np.ndarray:
    lib = load()
    st = np.zeros(134, np.uint8)
    lib.hevc_oracle_context_init(slice_qp, st.ctypes.data)
    return st


def test_binarization(kind: int, arg: int, bins) -> tuple[int, int]:
    """(value, bins consumed) of one binarisation fed from a bin list (kinds: see hevc_oracle.c)."""
    lib = load()
    b = bytes(int(bool(x)) for x in bins)
    used = C.c_int()
    v = lib.hevc_oracle_test_binarization(kind, arg, b, len(b), C.byref(used))
    return int(v), used.value
