"""ctypes binding of oracle/liboracle.so — the CPU checker (test infrastructure).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this."""
import ctypes as C
import os
import subprocess

import numpy as np

from jxl_rs_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


class JxoTaps(C.Structure):
    _fields_ = [("coeffs", C.c_void_p), ("xyb_idct", C.c_void_p), ("xyb_filtered", C.c_void_p)]


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = C.CDLL(path)
    lib.jxo_last_error.restype = C.c_char_p
    lib.jxo_decode_file.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(JxoTaps), C.c_int]
    lib.jxo_file_info.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(abi.JxgImageInfo)]
    lib.jxo_decode_frame.argtypes = [C.POINTER(abi.JxgFrameDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                     C.c_void_p, C.c_size_t, C.POINTER(JxoTaps), C.c_int, C.POINTER(C.c_uint32)]
    lib.jxo_idct2d.argtypes = [C.c_int, C.c_int, C.c_void_p]
    lib.jxo_reinterpreting_dct2d.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    lib.jxo_transform_to_pixels.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    lib.jxo_gaborish.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float]
    lib.jxo_xyb_to_linear.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]
    lib.jxo_linear_to_srgb.argtypes = [C.c_int, C.c_void_p]
    _LIB = lib
    return lib


def file_info(data: bytes):
    lib = load()
    info = abi.JxgImageInfo()
    r = lib.jxo_file_info(data, len(data), C.byref(info))
    if r != 0:
        raise abi.JxgError(r, lib.jxo_last_error().decode())
    return info


def decode_file(data: bytes, fmt=abi.FORMAT_RGB_U8, taps=False, threads=0):
    """Returns (pixels, taps dict). pixels: HxWx3 u8 / HxWx4 u8 / HxWx3 f32 / HxWx3 u16 / HxWx3 f16."""
    lib = load()
    info = file_info(data)
    w, h = info.width, info.height              # display orientation
    cw, ch = info.coded_width, info.coded_height  # as coded: parity taps and the XYB debug format
    if fmt == abi.FORMAT_RGB_F32:
        out = np.zeros((h, w, 3), np.float32)
    elif fmt == abi.FORMAT_RGBA_U8:
        out = np.zeros((h, w, 4), np.uint8)
    elif fmt == abi.FORMAT_XYB_F32_PLANAR:
        out = np.zeros((3, ch, cw), np.float32)
    elif fmt == abi.FORMAT_RGB_U16:
        out = np.zeros((h, w, 3), np.uint16)
    elif fmt == abi.FORMAT_RGB_F16:
        out = np.zeros((h, w, 3), np.float16)
    else:
        out = np.zeros((h, w, 3), np.uint8)
    stride = out.strides[0] if fmt != abi.FORMAT_XYB_F32_PLANAR else out.strides[1]
    t = None
    tap_arrays = {}
    if taps:
        ps, pr = (cw + 7) // 8 * 8, (ch + 7) // 8 * 8
        tap_arrays["coeffs"] = np.zeros((info.num_groups, 3, 65536), np.int32)
        tap_arrays["xyb_idct"] = np.zeros((3, pr, ps), np.float32)
        tap_arrays["xyb_filtered"] = np.zeros((3, ch, cw), np.float32)
        t = JxoTaps(tap_arrays["coeffs"].ctypes.data, tap_arrays["xyb_idct"].ctypes.data, tap_arrays["xyb_filtered"].ctypes.data)
    r = lib.jxo_decode_file(data, len(data), fmt, out.ctypes.data, stride, C.byref(t) if t else None, threads)
    if r != 0:
        raise abi.JxgError(r, lib.jxo_last_error().decode())
    return out, tap_arrays


def _load_modular():
    lib = load()
    if not getattr(lib, "_modular_ready", False):
        lib.jxo_modular_last_error.restype = C.c_char_p
        lib.jxo_modular_info.argtypes = [C.c_char_p, C.c_size_t] + [C.POINTER(C.c_uint32)] * 4
        lib.jxo_modular_orientation.argtypes = [C.c_char_p, C.c_size_t]
        lib.jxo_decode_modular_file.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        lib._modular_ready = True
    return lib


def modular_info(data: bytes):
    """(width, height, colour channels, groups) of a Modular-encoded file."""
    lib = _load_modular()
    v = [C.c_uint32() for _ in range(4)]
    r = lib.jxo_modular_info(data, len(data), *v)
    if r != 0:
        raise abi.JxgError(r, lib.jxo_modular_last_error().decode())
    return tuple(x.value for x in v)


def decode_modular_file(data: bytes, planes=False):
    """CPU decode of a Modular frame: H x W x 3 u8 (and the 3 i32 planes before the u8 conversion)."""
    lib = _load_modular()
    w, h, _, _ = modular_info(data)
    out = np.zeros((h, w, 3), np.uint8)
    transposed = lib.jxo_modular_orientation(data, len(data)) >= 5
    pl = (np.zeros((3, w, h), np.int32) if transposed else np.zeros((3, h, w), np.int32)) if planes else None  # as coded
    r = lib.jxo_decode_modular_file(data, len(data), out.ctypes.data, out.strides[0], pl.ctypes.data if planes else None)
    if r != 0:
        raise abi.JxgError(r, lib.jxo_modular_last_error().decode())
    return (out, pl) if planes else out
