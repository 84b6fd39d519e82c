"""GPU (libjxgpu.so, through the C ABI) vs the CPU oracle on the same bitstreams.

Bars: AC coefficients bit-exact; XYB planes after IDCT and after the
loop filters within 1e-3 abs-or-rel (the reference's own SIMD-vs-scalar allowance,
jxl/src/render/stages/xyb.rs:355); RGB u8 output within 1 LSB.
"""
import glob
import os

import numpy as np
import pytest

from jxl_rs_b200 import abi

pytestmark = pytest.mark.gpu

FILES = ["zoltan_tasi_unsplash.jxl", "green_queen_vardct_e3.jxl", "progressive_ac.jxl", "has_permutation.jxl",
         "opsin_inverse.jxl", "3x3_srgb_lossy.jxl", "basic.jxl", "lossy_with_icc.jxl", "grayscale.jxl"]


@pytest.fixture(scope="module")
def ctx():
    import jxl_rs_b200 as j
    c = j.JxgContext(0)
    yield c
    c.close()


def close_abs_rel(a, b, tol):
    d = np.abs(a - b)
    return np.all((d <= tol) | (d <= tol * np.maximum(np.abs(a), np.abs(b))))


@pytest.mark.parametrize("name", FILES)
def test_real_file_parity(ctx, golden_dir, name):
    import torch
    import jxl_rs_b200 as j
    from tests import oracle_binding as ob
    data = open(os.path.join(golden_dir, "jxl", name), "rb").read()
    ref_u8, taps = ob.decode_file(data, abi.FORMAT_RGB_U8, taps=True)
    fr = j.ParsedFrame(data)
    # 1) coefficients + IDCT planes (stop after K2)
    out = torch.empty((fr.height, fr.width, 3), dtype=torch.uint8).pin_memory()
    b = j.Batch(ctx, 1)
    b.add(fr, out.data_ptr(), fr.width * 3, abi.FORMAT_RGB_U8, False)
    b.set_debug_stop(2)
    b.run()
    b.wait()
    co = b.read_coeffs(0)
    assert np.array_equal(co, taps["coeffs"]), "AC coefficients are not bit-exact"
    xyb0 = b.read_xyb(0, 0)
    assert close_abs_rel(xyb0, taps["xyb_idct"], 1e-3), f"IDCT planes differ: max {np.abs(xyb0 - taps['xyb_idct']).max()}"
    b.close()
    # 2) full pipeline
    b = j.Batch(ctx, 1)
    b.add(fr, out.data_ptr(), fr.width * 3, abi.FORMAT_RGB_U8, False)
    b.run()
    b.wait()
    got = out.numpy()
    diff = np.abs(got.astype(np.int32) - ref_u8.astype(np.int32))
    assert diff.max() <= 1, f"u8 output differs by {diff.max()} LSB"
    assert (diff > 0).mean() < 0.01
    st = b.stats()
    assert st["kernel_launches"] >= 3
    b.close()
    # 3) XYB planes after Gaborish + EPF (tap format)
    xout = torch.empty((3, fr.height, fr.width), dtype=torch.float32).pin_memory()
    b = j.Batch(ctx, 1)
    b.add(fr, xout.data_ptr(), fr.width * 4, abi.FORMAT_XYB_F32_PLANAR, False)
    b.run()
    b.wait()
    b.close()
    assert close_abs_rel(xout.numpy(), taps["xyb_filtered"], 1e-3), f"filtered planes differ: max {np.abs(xout.numpy() - taps['xyb_filtered']).max()}"


def test_batch_of_mixed_frames(ctx, golden_dir):
    """Several different frames in one batch (different sizes, filters, pass counts)."""
    import jxl_rs_b200 as j
    from tests import oracle_binding as ob
    names = ["green_queen_vardct_e3.jxl", "zoltan_tasi_unsplash.jxl", "progressive_ac.jxl"]
    datas = [open(os.path.join(golden_dir, "jxl", n), "rb").read() for n in names]
    outs = j.decode_files(ctx, datas)
    for d, o in zip(datas, outs):
        ref, _ = ob.decode_file(d)
        assert np.abs(o.numpy().astype(np.int32) - ref.astype(np.int32)).max() <= 1


def test_f32_and_rgba_outputs(ctx, golden_dir):
    import jxl_rs_b200 as j
    from tests import oracle_binding as ob
    data = open(os.path.join(golden_dir, "jxl", "green_queen_vardct_e3.jxl"), "rb").read()
    f32 = j.decode_files(ctx, [data], j.JxlPixelFormat("RGB", "F32"))[0].numpy()
    ref, _ = ob.decode_file(data, abi.FORMAT_RGB_F32)
    assert close_abs_rel(f32, ref, 1e-3)
    rgba = j.decode_files(ctx, [data], j.JxlPixelFormat("RGBA", "U8"))[0].numpy()
    ref4, _ = ob.decode_file(data, abi.FORMAT_RGBA_U8)
    assert np.abs(rgba.astype(np.int32) - ref4.astype(np.int32)).max() <= 1
    assert (rgba[..., 3] == 255).all()


def test_corrupt_stream_reports_error(ctx, golden_dir):
    """Flipping bits inside an HF section must surface as an entropy error, not a crash."""
    import jxl_rs_b200 as j
    data = bytearray(open(os.path.join(golden_dir, "jxl", "green_queen_vardct_e3.jxl"), "rb").read())
    for i in range(len(data) - 3000, len(data) - 2000):
        data[i] ^= 0x5a
    with pytest.raises(abi.JxgError):
        j.decode_files(ctx, [bytes(data)])
