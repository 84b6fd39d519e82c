"""Pixel parity against the ACTUAL jxl-rs binary, for whoever has one: tools/make_reference_goldens.sh writes
float32 .npy decodes of jxl_cli (jxl_cli/src/enc/numpy.rs) into tests/golden/pixels/. The build image of this
repository has no Rust toolchain, so the directory is normally empty and these tests skip; with goldens present they
pin the CPU oracle (here) and the CUDA path (`-m gpu`) to the reference's own output.
Tolerance: max-abs 1e-3 on sRGB-encoded float samples in [0, 1] (the reference's SIMD-vs-scalar
allowance xyb.rs:355; one 8-bit LSB is 3.9e-3)."""
import glob
import os

import numpy as np
import pytest

from jxl_rs_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PIXELS = os.path.join(ROOT, "tests", "golden", "pixels")
GOLDENS = sorted(glob.glob(os.path.join(PIXELS, "*.npy")))
TOL = 1e-3


def _input_for(npy):
    name = os.path.basename(npy)[:-4]
    for d in (os.path.join(ROOT, "tests", "golden", "jxl"), PIXELS):
        p = os.path.join(d, name + ".jxl")
        if os.path.exists(p):
            return p
    return None


def _compare(ours, golden):
    """`ours`: F32 output of the path. Since the decoder follows the reference's output-profile rule (a non-ICC embedded
    encoding is the output encoding for float samples too, api/inner/codestream_parser/image_info.rs:204-237), the two
    arrays are in the same encoding and are compared directly: one expected answer, no alternatives."""
    g = np.load(golden) if isinstance(golden, str) else golden
    assert g.ndim == 4 and g.shape[0] >= 1, f"unexpected golden shape {g.shape}"
    g = g[0][..., :3]
    if g.shape[-1] == 1:  # grey output of the reference: our path writes R = G = B
        g = np.repeat(g, 3, axis=-1)
    assert g.shape == ours.shape, f"size mismatch {g.shape} vs {ours.shape}"
    err = float(np.abs(ours - g).max())
    assert err <= TOL, f"max abs error vs jxl_cli: {err:.2e}"


@pytest.mark.skipif(not GOLDENS, reason="no jxl_cli goldens (tools/make_reference_goldens.sh needs a Rust toolchain)")
@pytest.mark.parametrize("golden", GOLDENS, ids=[os.path.basename(g) for g in GOLDENS])
def test_oracle_matches_jxl_cli(golden):
    from tests import oracle_binding as ob
    src = _input_for(golden)
    if src is None:
        pytest.skip("input bitstream not in the repository")
    data = open(src, "rb").read()
    try:
        out, _ = ob.decode_file(data, abi.FORMAT_RGB_F32)
    except abi.JxgError as e:
        pytest.skip(f"outside the hot-path scope: {e}")
    _compare(out, golden)


@pytest.mark.gpu
@pytest.mark.skipif(not GOLDENS, reason="no jxl_cli goldens (tools/make_reference_goldens.sh needs a Rust toolchain)")
@pytest.mark.parametrize("golden", GOLDENS, ids=[os.path.basename(g) for g in GOLDENS])
def test_cuda_path_matches_jxl_cli(golden):
    import jxl_rs_b200 as j
    src = _input_for(golden)
    if src is None:
        pytest.skip("input bitstream not in the repository")
    data = open(src, "rb").read()
    ctx = j.JxgContext(0)
    try:
        try:
            out = j.decode_files(ctx, [data], j.JxlPixelFormat("RGB", "F32"))[0]
        except abi.JxgError as e:
            pytest.skip(f"outside the hot-path scope: {e}")
        _compare(out.numpy(), golden)
    finally:
        ctx.close()


def test_comparison_logic_on_a_simulated_golden():
    """The harness itself: a 'golden' made from the oracle's own output, stored the way jxl_cli stores it
    (frames x H x W x C, in the image's output encoding), must compare clean, and a perturbed one must not."""
    import synth
    from tests import oracle_binding as ob
    data = synth.encode_synthetic(96, 64, 3, 0.5, 2, 1, 1)
    out, _ = ob.decode_file(data, abi.FORMAT_RGB_F32)
    fake = out[None].copy()
    _compare(out, fake)
    with pytest.raises(AssertionError):
        _compare(out, fake + np.float32(0.01))
