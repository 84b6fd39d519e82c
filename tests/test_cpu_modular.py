"""CPU checks of the Modular path: the oracle against known answers and lossless round trips,
the host front-end's stream plan, and the synthetic writer. No GPU."""
import os

import numpy as np
import pytest


def test_known_answer_3x3(golden_dir):
    """3x3_srgb_lossless.jxl (reference fixture): primaries, mid tones, white / grey / black."""
    from tests import oracle_binding as ob
    data = open(os.path.join(golden_dir, "jxl", "3x3_srgb_lossless.jxl"), "rb").read()
    out = ob.decode_modular_file(data).reshape(-1, 3)
    want = [[255, 0, 0], [0, 255, 0], [0, 0, 255], [128, 64, 64], [64, 128, 64], [64, 64, 128], [255, 255, 255],
            [128, 128, 128], [0, 0, 0]]
    assert out.tolist() == want


@pytest.mark.parametrize("name,shape", [("green_queen_modular_e3.jxl", (589, 438)), ("lz77_flower.jxl", (244, 834)),
                                        ("delta_palette.jxl", (751, 555)), ("grayscale_public_university.jxl", (1620, 2880)),
                                        ("tree_max_property_20.jxl", (1024, 1024))])
def test_real_modular_fixtures_self_verify(golden_dir, name, shape):
    """Every ANS stream of the reference's Modular fixtures must end in its checksum state without over-read
    (decode.rs:400) — through RCT, palette, Squeeze, LZ77, the weighted predictor and reference-channel properties."""
    from tests import oracle_binding as ob
    data = open(os.path.join(golden_dir, "jxl", name), "rb").read()
    out = ob.decode_modular_file(data)
    assert out.shape == (shape[0], shape[1], 3)
    assert out.std() > 1.0  # not a blank image


def test_lz77_flower_is_three_copies(golden_dir):
    """lz77_flower.jxl is the same 278-pixel-wide picture three times side by side: a content check of the LZ77 path."""
    from tests import oracle_binding as ob
    out = ob.decode_modular_file(open(os.path.join(golden_dir, "jxl", "lz77_flower.jxl"), "rb").read())
    assert np.array_equal(out[:, :278], out[:, 278:556]) and np.array_equal(out[:, :278], out[:, 556:834])


@pytest.mark.parametrize("w,h,rct,sq,tk", [(64, 48, 0, 0, 0), (300, 200, 6, 0, 1), (300, 200, 0, 1, 2), (700, 530, 6, 1, 1),
                                           (257, 513, 6, 1, 2), (1, 1, 0, 0, 0), (9, 1000, 6, 1, 0), (300, 260, 6, 0, 3),
                                           (513, 300, 0, 1, 3)])
def test_lossless_roundtrip(w, h, rct, sq, tk):
    """Writer (forward RCT / Squeeze written from the decoder definitions) -> oracle == source image, bit-exact."""
    import synth
    from tests import oracle_binding as ob
    data = synth.encode_modular(w, h, 11, rct, sq, tk)
    assert np.array_equal(ob.decode_modular_file(data), synth.modular_source(w, h, 11))


@pytest.mark.parametrize("w,h,tk", [(300, 260, 1), (64, 48, 0), (700, 530, 2), (1030, 600, 3)])
def test_palette_roundtrip_uses_explicit_and_implicit_entries(w, h, tk):
    """Writer's palette variant (explicit entries + the 4x4x4 and 5x5x5 implicit cubes, palette.rs:17-163) -> oracle ==
    the snapped source picture; the index channel really holds indices of all three kinds."""
    import synth
    from tests import oracle_binding as ob
    data = synth.encode_modular(w, h, 12, 0, 0, tk, palette=1)
    src = synth.modular_source(w, h, 12, palette=1)
    assert np.array_equal(ob.decode_modular_file(data), src)
    left, right = src[:, :w // 4].reshape(-1), src[:, w // 4:].reshape(-1)
    assert set(np.unique(left)) <= {32, 95, 159, 223} and set(np.unique(right)) <= {0, 63, 127, 191, 255}


def test_roundtrip_of_extreme_values():
    """Caller-supplied image with saturated checkerboards and flat areas (largest residuals, zero residuals)."""
    import synth
    from tests import oracle_binding as ob
    rng = np.random.default_rng(5)
    img = np.zeros((300, 520, 3), np.uint8)
    img[::2, ::2] = 255
    img[100:200, 100:400] = rng.integers(0, 256, (100, 300, 3), dtype=np.uint8)
    img[250:] = 17
    for sq in (0, 1):
        data = synth.encode_modular(520, 300, 0, 6, sq, 2, source=img)
        assert np.array_equal(ob.decode_modular_file(data), img)


def test_corrupt_stream_is_rejected():
    import synth
    from jxl_rs_b200 import abi
    from tests import oracle_binding as ob
    data = bytearray(synth.encode_modular(600, 520, 9, 6, 0, 1))
    data[-2000] ^= 0x5A
    with pytest.raises(abi.JxgError):
        ob.decode_modular_file(bytes(data))
