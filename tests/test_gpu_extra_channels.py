"""GPU parity on VarDCT frames that carry extra channels (alpha): the device path decodes their colour channels
(the extra channels' Modular streams are stepped over by the front-end; in the HF sections they sit behind the AC
coefficients the entropy kernel reads)."""
import os

import numpy as np
import pytest

from jxl_rs_b200 import abi

pytestmark = pytest.mark.gpu

FILES = ["3x3a_srgb_lossy.jxl", "alpha_premultiplied.jxl", "dice.jxl", "squeeze_alpha.jxl", "upsampled_alpha.jxl"]


@pytest.mark.parametrize("name", FILES)
def test_colour_of_frames_with_alpha(golden_dir, name):
    import jxl_rs_b200 as j
    from tests import oracle_binding as ob
    data = open(os.path.join(golden_dir, "jxl", name), "rb").read()
    ref, taps = ob.decode_file(data, abi.FORMAT_RGB_U8, taps=True)
    ctx = j.JxgContext(0)
    try:
        frame = j.ParsedFrame(data)
        import torch
        out = torch.empty((frame.height, frame.width, 3), dtype=torch.uint8).pin_memory()
        b = j.Batch(ctx, 1)
        b.add(frame, out.data_ptr(), frame.width * 3, abi.FORMAT_RGB_U8, False)
        b.run()
        b.wait()
        assert np.array_equal(b.read_coeffs(0), taps["coeffs"]), "coefficients differ from the oracle"
        d = np.abs(out.numpy().astype(np.int32) - ref.astype(np.int32)).max()
        assert d <= 1, f"u8 output differs from the oracle by {d} LSB"
        b.close()
    finally:
        ctx.close()
