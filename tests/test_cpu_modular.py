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


def _random_tree(rng, prop, depth, allow_other, split_lo, split_hi, n_ctx):
    """Random MA tree as a node list {property, value, left | predictor, right | multiplier, ctx}: splits on the channel
    (0), the stream id (1) and `prop`; optionally one split on another property."""
    nodes = []

    def build(d):
        i = len(nodes)
        nodes.append(None)
        if d == 0 or rng.random() < 0.25:
            plain = rng.random() < 0.6
            nodes[i] = [-1, 0 if plain else int(rng.integers(-5, 6)), int(rng.integers(0, 14)), 1 if plain else int(rng.integers(1, 4)),
                        int(rng.integers(0, n_ctx))]
            return i
        r = rng.random()
        if r < 0.15:
            p, v = 0, int(rng.integers(0, 3))
        elif r < 0.25:
            p, v = 1, int(rng.integers(20, 30))
        elif allow_other and r < 0.32:
            p, v = (prop % 13) + 2 if (prop % 13) + 2 != prop else 3, int(rng.integers(-50, 50))
        else:
            p, v = prop, int(rng.integers(split_lo, split_hi))
        nodes[i] = [p, v, 0, 0, 0]
        nodes[i][2] = build(d - 1)
        nodes[i][3] = build(d - 1)
        return i

    build(depth)
    return nodes


def _walk(nodes, channel, stream, prop, v):
    i = 0
    while nodes[i][0] >= 0:
        p, val = nodes[i][0], nodes[i][1]
        x = channel if p == 0 else (stream if p == 1 else v)
        assert p in (0, 1, prop)
        i = nodes[i][2] if x > val else nodes[i][3]
    return i


def test_walk_tables_agree_with_the_tree_walk():
    """The host logic behind the device's table walk (jxg_modular_walk_table = build_walk_table of the Modular batch
    engine): for random trees that split on one property below any channel / stream decisions, every property value —
    far outside the table's [-1024, 1023] included — must land on the leaf the node walk (tree.rs:360-390) reaches, with that
    leaf's predictor, cluster and plain flag; trees that split on a second property or beyond the clampable range must be
    refused (generic walk)."""
    import ctypes as C
    from jxl_rs_b200 import abi
    lib = abi.load_library()
    rng = np.random.default_rng(2024)
    accepted = refused = 0
    for trial in range(400):
        prop = int(rng.integers(2, 16))
        allow_other = trial % 5 == 4
        wide = trial % 7 == 6  # split values beyond what clamping preserves
        n_ctx = 40
        nodes = _random_tree(rng, prop, int(rng.integers(0, 6)), allow_other, -1500 if wide else -1024, 1500 if wide else 1023, n_ctx)
        cmap = rng.integers(0, 9, n_ctx).astype(np.uint8)
        flat = np.array(nodes, np.int32).reshape(-1)
        channel, stream = int(rng.integers(0, 4)), int(rng.integers(18, 32))
        lut = np.zeros(2048, np.uint32)
        p_out = C.c_uint32(0)
        r = lib.jxg_modular_walk_table(flat.ctypes.data_as(C.POINTER(C.c_int32)), len(nodes), cmap.ctypes.data_as(C.POINTER(C.c_uint8)),
                                       n_ctx, channel, stream, lut.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(p_out))
        assert r in (0, 1), r
        # what the walk can reach for this channel / stream
        reach, stack = [], [0]
        while stack:
            i = stack.pop()
            reach.append(i)
            p, val = nodes[i][0], nodes[i][1]
            if p < 0:
                continue
            if p in (0, 1):
                x = channel if p == 0 else stream
                stack.append(nodes[i][2] if x > val else nodes[i][3])
            else:
                stack += [nodes[i][2], nodes[i][3]]
        props = {nodes[i][0] for i in reach if nodes[i][0] >= 2}
        if r == 0:
            refused += 1
            out_of_range = any(nodes[i][0] >= 2 and not (-1024 <= nodes[i][1] <= 1022) for i in reach)
            assert len(props) > 1 or out_of_range, (trial, props)
            continue
        accepted += 1
        assert len(props) <= 1
        the_prop = next(iter(props)) if props else None  # may be the "other" property when only that one is reachable
        assert p_out.value == (the_prop if props else 0xff)
        for v in list(rng.integers(-4000, 4000, 300)) + [-1025, -1024, -1023, 1022, 1023, 1024, 0]:
            leaf = _walk(nodes, channel, stream, the_prop, int(v))
            e = int(lut[min(max(int(v), -1024), 1023) + 1024])
            n = nodes[leaf]
            assert e >> 16 == leaf, (trial, v)
            assert e & 15 == n[2] & 15 and (e >> 4) & 255 == int(cmap[n[4]])
            assert bool(e & (1 << 12)) == (n[1] == 0 and n[3] == 1)
    assert accepted > 150 and refused >= 15, (accepted, refused)
