"""FFmpeg HEVC decoder as an independent bit-exact oracle.  TEST INFRASTRUCTURE ONLY.

The reference crate cannot reconstruct a pixel (src/hevc/slice.rs:249-255 are todo!()), and libheif /
libde265 are absent, so full-reconstruction parity is pinned against the native `hevc` decoder of the
FFmpeg build bundled with opencv-python-headless (libavcodec.so.62), driven through ctypes
(SURVEY.md section 8(c)).  Struct offsets below are for libavcodec 62 / libavutil 60 and are checked at
load time against known-answer planes (tests/golden/fixture_hashes.json).

Only tests/, golden-vector generators and bench.py's CPU-baseline leg import this module.
"""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np

_AV_CODEC_ID_HEVC = 173
_PKT_DATA, _PKT_SIZE = 24, 32
_FR_DATA, _FR_LINESIZE, _FR_WIDTH, _FR_HEIGHT, _FR_FORMAT = 0, 64, 104, 108, 116

_libs = None


def _load():
    global _libs
    if _libs is not None:
        return _libs
    import cv2  # noqa: F401  (pre-loads libdrm & friends the bundled FFmpeg links against)

    d = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
    avutil = glob.glob(os.path.join(d, "libavutil-*.so.*"))
    avcodec = glob.glob(os.path.join(d, "libavcodec-*.so.*"))
    if not avutil or not avcodec:
        raise ImportError("bundled FFmpeg (opencv_python_headless.libs) not found")
    for dep in ("libswresample-*.so.*",):
        for p in glob.glob(os.path.join(d, dep)):
            try:
                C.CDLL(p, mode=C.RTLD_GLOBAL)
            except OSError:
                pass
    u = C.CDLL(avutil[0], mode=C.RTLD_GLOBAL)
    c = C.CDLL(avcodec[0], mode=C.RTLD_GLOBAL)
    c.avcodec_find_decoder.restype = C.c_void_p
    c.avcodec_find_decoder.argtypes = [C.c_int]
    c.avcodec_alloc_context3.restype = C.c_void_p
    c.avcodec_alloc_context3.argtypes = [C.c_void_p]
    c.avcodec_open2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    c.avcodec_free_context.argtypes = [C.POINTER(C.c_void_p)]
    c.av_packet_alloc.restype = C.c_void_p
    c.av_packet_free.argtypes = [C.POINTER(C.c_void_p)]
    c.av_new_packet.argtypes = [C.c_void_p, C.c_int]
    c.av_packet_unref.argtypes = [C.c_void_p]
    c.avcodec_send_packet.argtypes = [C.c_void_p, C.c_void_p]
    c.avcodec_receive_frame.argtypes = [C.c_void_p, C.c_void_p]
    c.avcodec_flush_buffers.argtypes = [C.c_void_p]
    u.av_frame_alloc.restype = C.c_void_p
    u.av_frame_free.argtypes = [C.POINTER(C.c_void_p)]
    u.av_frame_unref.argtypes = [C.c_void_p]
    _libs = (u, c)
    return _libs


def annexb(nals) -> bytes:
    """Concatenate NAL units (bytes incl. 2-byte header, escaped) with 4-byte start codes."""
    return b"".join(b"\x00\x00\x00\x01" + bytes(n) for n in nals)


class FFmpegHevc:
    """One decoder context; decode_picture() takes an Annex-B access unit with VPS/SPS/PPS in band."""

    def __init__(self, threads: int = 1):
        u, c = _load()
        self.u, self.c = u, c
        codec = c.avcodec_find_decoder(_AV_CODEC_ID_HEVC)
        if not codec:
            raise RuntimeError("FFmpeg build has no hevc decoder")
        self.ctx = C.c_void_p(c.avcodec_alloc_context3(codec))
        if c.avcodec_open2(self.ctx, codec, None) < 0:
            raise RuntimeError("avcodec_open2 failed")
        self.pkt = C.c_void_p(c.av_packet_alloc())
        self.frame = C.c_void_p(u.av_frame_alloc())

    def close(self):
        if self.ctx:
            self.u.av_frame_free(C.byref(self.frame))
            self.c.av_packet_free(C.byref(self.pkt))
            self.c.avcodec_free_context(C.byref(self.ctx))
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode_picture(self, au: bytes):
        """-> [Y, Cb, Cr] uint8 arrays (Cb/Cr absent for 4:0:0)."""
        c, u = self.c, self.u
        if c.av_new_packet(self.pkt, len(au)) < 0:
            raise MemoryError
        data_ptr = C.c_void_p.from_address(self.pkt.value + _PKT_DATA).value
        C.memmove(data_ptr, au, len(au))
        rc = c.avcodec_send_packet(self.ctx, self.pkt)
        c.av_packet_unref(self.pkt)
        if rc < 0:
            c.avcodec_flush_buffers(self.ctx)
            raise RuntimeError(f"avcodec_send_packet failed: {rc}")
        c.avcodec_send_packet(self.ctx, None)  # drain
        rc = c.avcodec_receive_frame(self.ctx, self.frame)
        if rc < 0:
            c.avcodec_flush_buffers(self.ctx)
            raise RuntimeError(f"avcodec_receive_frame failed: {rc}")
        base = self.frame.value
        w = C.c_int.from_address(base + _FR_WIDTH).value
        h = C.c_int.from_address(base + _FR_HEIGHT).value
        fmt = C.c_int.from_address(base + _FR_FORMAT).value
        # 0 yuv420p, 12 yuvj420p, 8 gray8
        if fmt not in (0, 12, 8):
            raise RuntimeError(f"unexpected pixel format {fmt}")
        planes = []
        for i in range(1 if fmt == 8 else 3):
            ptr = C.c_void_p.from_address(base + _FR_DATA + 8 * i).value
            ls = C.c_int.from_address(base + _FR_LINESIZE + 4 * i).value
            pw, ph = (w, h) if i == 0 else ((w + 1) // 2, (h + 1) // 2)
            buf = (C.c_uint8 * (ls * ph)).from_address(ptr)
            a = np.frombuffer(buf, dtype=np.uint8).reshape(ph, ls)[:, :pw].copy()
            planes.append(a)
        u.av_frame_unref(self.frame)
        c.avcodec_flush_buffers(self.ctx)
        return planes
