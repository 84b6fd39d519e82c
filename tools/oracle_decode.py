#!/usr/bin/env python3
"""Dev tool: decode .jxl files with the CPU oracle and write PNGs."""
import ctypes, sys, os, numpy as np, time
from PIL import Image
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle', 'liboracle.so'))
class Info(ctypes.Structure):
    _fields_ = [('width', ctypes.c_uint32), ('height', ctypes.c_uint32), ('num_groups', ctypes.c_uint32),
                ('num_passes', ctypes.c_uint32), ('encoding', ctypes.c_uint32), ('hf_bytes', ctypes.c_uint64)]
lib.jxo_last_error.restype = ctypes.c_char_p
for path in sys.argv[1:]:
    data = open(path, 'rb').read()
    info = Info()
    r = lib.jxo_file_info(data, len(data), ctypes.byref(info))
    if r != 0:
        print(path, 'info error', r, lib.jxo_last_error().decode()); continue
    out = np.zeros((info.height, info.width, 3), np.uint8)
    t = time.time()
    r = lib.jxo_decode_file(data, len(data), 0, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(info.width * 3), None, 8)
    dt = time.time() - t
    print(path, info.width, info.height, 'ret', r, lib.jxo_last_error().decode() if r else '', '%.1f MP/s' % (info.width * info.height / dt / 1e6))
    if r == 0:
        Image.fromarray(out).save('/tmp/' + os.path.basename(path) + '.png')
