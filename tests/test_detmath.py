"""rr_detmath.h (IEEE-only elementary functions shared by the kernels and the oracle) against libm/numpy,
and Philox4x32-10 against the Random123 known-answer vectors."""
import ctypes as C

import numpy as np


def _ulp_err(got, ref):
    ref = np.asarray(ref)
    spacing = np.abs(np.spacing(ref))
    with np.errstate(invalid="ignore"):
        e = np.abs(got.astype(np.float64) - ref.astype(np.float64)) / spacing.astype(np.float64)
    e[np.isnan(got) & np.isnan(ref)] = 0
    return np.nanmax(e)


def test_double_functions_within_ulps_of_libm(oracle_mod):
    rng = np.random.default_rng(1)
    x = rng.uniform(-7.0, 7.0, 400000)
    assert _ulp_err(oracle_mod.detmath_f64(0, x), np.sin(x)) <= 2.0
    assert _ulp_err(oracle_mod.detmath_f64(1, x), np.cos(x)) <= 2.0
    assert _ulp_err(oracle_mod.detmath_f64(2, x), np.tan(x)) <= 4.0
    y = rng.uniform(-1.0, 1.0, 400000)
    assert _ulp_err(oracle_mod.detmath_f64(3, y), np.arcsin(y)) <= 3.0
    assert _ulp_err(oracle_mod.detmath_f64(4, y), np.arccos(y)) <= 2.0
    e = rng.uniform(-90.0, 5.0, 400000)
    assert _ulp_err(oracle_mod.detmath_f64(5, e), np.exp(e)) <= 2.0
    p = np.abs(x) + 1e-300
    assert _ulp_err(oracle_mod.detmath_f64(6, p), np.log(p)) <= 2.0


def test_float_functions_within_one_ulp(oracle_mod):
    rng = np.random.default_rng(2)
    y = rng.uniform(-1.0, 1.0, 300000).astype(np.float32)
    assert _ulp_err(oracle_mod.detmath_f32(0, y), np.arccos(y.astype(np.float64)).astype(np.float32)) <= 1.0
    x = rng.uniform(-7.0, 7.0, 300000).astype(np.float32)
    assert _ulp_err(oracle_mod.detmath_f32(1, x), np.cos(x.astype(np.float64)).astype(np.float32)) <= 1.0
    e = rng.uniform(-80.0, 3.0, 300000).astype(np.float32)
    assert _ulp_err(oracle_mod.detmath_f32(2, e), np.exp(e.astype(np.float64)).astype(np.float32)) <= 1.0
    c = np.abs(y)
    for expo in (1.0, 8.0, 40.0, 100.0, 3000.0):
        ref = np.power(c.astype(np.float64), expo).astype(np.float32)
        assert _ulp_err(oracle_mod.detmath_f32(3, c, expo), ref) <= 1.0
    ref4 = np.power(c.astype(np.float64), 4.0).astype(np.float32)
    assert np.array_equal(oracle_mod.detmath_f32(4, c), ref4)          # pow(x, 4.0) is exact-rounded


def test_special_cases(oracle_mod):
    f = oracle_mod.detmath_f32
    assert np.isnan(f(0, np.array([1.0000001, -1.0000001], np.float32))).all()      # acos(|x|>1) = NaN (reference quirk)
    assert f(0, np.array([1.0], np.float32))[0] == 0.0
    assert f(3, np.array([-0.5], np.float32), 3.0)[0] == -0.125                       # odd integer exponent
    assert f(3, np.array([-0.5], np.float32), 2.0)[0] == 0.25
    assert np.isnan(f(3, np.array([-0.5], np.float32), 2.5)[0])
    assert f(3, np.array([0.0], np.float32), 3000.0)[0] == 0.0
    assert f(3, np.array([-4.37e-8], np.float32), 3000.0)[0] == 0.0                   # cos(pi/2 as float) ** 3000
    assert f(3, np.array([0.7], np.float32), 0.0)[0] == 1.0
    assert oracle_mod.detmath_f64(5, np.array([-1000.0]))[0] == 0.0


def test_philox_known_answers(oracle_mod):
    L = oracle_mod.lib()

    def philox(ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        L.orc_philox(c, k, o)
        return [int(v) for v in o]
    # Random123 kat_vectors, philox4x32 with 10 rounds
    assert philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_noise_stream_is_keyed_and_uniform(oracle_mod):
    L = oracle_mod.lib()
    v = np.array([L.orc_noise_u01(7, 3, a, d) for a in range(20) for d in range(200)])
    assert (v >= 0).all() and (v < 1).all()
    assert abs(v.mean() - 0.5) < 0.02
    assert L.orc_noise_u01(7, 3, 5, 9) == L.orc_noise_u01(7, 3, 5, 9)
    assert L.orc_noise_u01(7, 3, 5, 9) != L.orc_noise_u01(7, 4, 5, 9)
    assert L.orc_noise_u01(7, 3, 5, 9) != L.orc_noise_u01(8, 3, 5, 9)
