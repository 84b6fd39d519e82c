"""Known-answer tests of the reference's entropy layer replayed on the host front-end
(which the oracle and the device path both consume). Sources: jxl/src/entropy_coding/ans.rs:463-485,
huffman.rs:516-527, hybrid_uint.rs:118-127, headers/permutation.rs:286-347, bit_reader.rs doctest :263-273,
frame/modular/predict.rs:564-593, frame/coeff_order.rs:155-178, frame/quant_weights.rs:1222-2139."""
import ctypes as C
import json
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def lib(oracle):
    return oracle


@pytest.fixture(scope="module")
def kat(golden_dir):
    return json.load(open(os.path.join(golden_dir, "kat.json")))


def ans_hist(lib, data, log_alpha):
    dist = (C.c_uint16 * 256)()
    single = C.c_int32(-2)
    r = lib.jxo_t_ans_histogram(bytes(data), len(data), log_alpha, dist, C.byref(single))
    return r, list(dist)[: 1 << log_alpha], single.value


def test_ans_single_symbol(lib):  # ans.rs:463-474
    r, dist, single = ans_hist(lib, [0b00100101, 0b01], 5)
    assert r == 0 and sum(dist) == 4096 and dist[20] == 4096 and single == 20
    r, _, _ = ans_hist(lib, [0b00101101, 0b000], 5)
    assert r != 0  # symbol beyond the alphabet


def test_ans_two_symbols(lib):  # ans.rs:476-485
    r, dist, single = ans_hist(lib, [0b10011111, 0b10010010, 0b00000000, 0b00010], 5)
    assert r == 0 and sum(dist) == 4096
    assert dist[10] == 256 and dist[20] == 4096 - 256 and single == -1


def test_prefix_byte_histogram(lib):  # huffman.rs:505-527
    hist = bytes([0b11101111, 0b00111111, 0, 1, 0, 0b10100000, 0b0110])
    expected = [8, 13, 21, 34, 55, 89, 144, 233]
    data = bytes(int(f"{v:08b}"[::-1], 2) for v in expected)
    out = (C.c_uint32 * 8)()
    assert lib.jxo_t_prefix_read(hist, len(hist), data, len(data), 8, out) == 0
    assert list(out) == expected


def test_hybrid_uint_invalid_config_does_not_crash(lib):  # hybrid_uint.rs:118-127
    data = bytes([10, 75, 10, 75, 168, 139, 132, 255, 244])
    v = C.c_uint32()
    lib.jxo_t_hybrid_decode_config(data, len(data), 1, 15, 1022, C.byref(v))


def test_bit_reader_lsb_first(lib):  # bit_reader.rs:262-273
    data = bytes([0x12, 0x34, 0x56, 0x78])
    nbits = (C.c_uint32 * 2)(1, 7)
    out = (C.c_uint64 * 2)()
    total = lib.jxo_t_bitreader(data, len(data), nbits, 2, out)
    assert total == 8 and out[1] == 0x12 >> 1
    nbits = (C.c_uint32 * 3)(4, 12, 16)
    out = (C.c_uint64 * 3)()
    lib.jxo_t_bitreader(data, len(data), nbits, 3, out)
    assert list(out) == [0x2, 0x341, 0x7856]


def test_lehmer(lib):  # permutation.rs:286-345
    def run(code, skip, size):
        c = (C.c_uint32 * len(code))(*code)
        out = (C.c_uint32 * size)()
        r = lib.jxo_t_lehmer(c, len(code), skip, size, out)
        return r, list(out)
    r, p = run([1, 1, 2, 3, 3, 6, 0, 1], 4, 16)
    assert r == 0 and p == [0, 1, 2, 3, 5, 6, 8, 10, 11, 15, 4, 9, 7, 12, 13, 14]
    r, p = run([2, 3, 0, 0, 0], 0, 5)
    assert r == 0 and p == [2, 4, 0, 1, 3]
    r, _ = run([4], 4, 8)
    assert r != 0


def test_weighted_predictor_golden(lib):  # predict.rs:564-593
    preds = (C.c_int64 * 4)()
    props = (C.c_int32 * 4)()
    lib.jxo_t_wp_golden(preds, props)
    assert list(zip(preds, props)) == [(135, 0), (110, -60), (165, 0), (153, -60)]


def test_natural_coeff_orders(lib, kat):  # coeff_order.rs:155-178
    for idx, name, n in [(0, "COEFF_ORDER_1X1", 64), (4, "COEFF_ORDER_2X1", 128)]:
        out = (C.c_uint32 * n)()
        assert lib.jxo_t_natural_order(idx, out, n) == n
        assert list(out) == kat[name]
    # every order is a permutation
    for idx, n in enumerate([64, 64, 256, 1024, 128, 256, 512, 4096, 2048, 16384, 8192, 65536, 32768]):
        out = (C.c_uint32 * n)()
        assert lib.jxo_t_natural_order(idx, out, n) == n
        assert sorted(out) == list(range(n))


def test_library_dequant_tables(lib, kat):  # quant_weights.rs:1222-2139 (tolerance 1e-5)
    cov_x = [1, 1, 1, 1, 2, 4, 1, 2, 1, 4, 2, 4, 1, 1, 1, 1, 1, 1, 8, 4, 8, 16, 8, 16, 32, 16, 32]
    cov_y = [1, 1, 1, 1, 2, 4, 2, 1, 4, 1, 4, 2, 1, 1, 1, 1, 1, 1, 8, 8, 4, 16, 16, 8, 32, 32, 16]
    target = kat["DEQUANT_TARGET_TABLE"]
    k = 0
    buf = (C.c_float * 65536)()
    for t in range(27):
        size = cov_x[t] * cov_y[t] * 64
        for c in range(3):
            assert lib.jxo_t_dequant_table(t, c, buf, 65536) == size
            for jj in range(0, size, size // 10):
                assert abs(buf[jj] - target[k]) < 1e-5, (t, c, jj)
                k += 1
    assert k == len(target)
