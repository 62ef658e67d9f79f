"""The two integer identities the fp32 consensus kernels count with (lsqrrecipes_b200/csrc/k_fast.cu), checked in numpy:

  * carry chain (count_carry): for the shifted residual s' = s + delta, `bits(s') < bits(2 delta)` as unsigned integers is
    exactly `0 <= s' < 2 delta` (with -0.0, negative values, NaN and infinities on the outlier side), and the carry out of
    `bits(s') + (2^32 - bits(2 delta))` is the outlier flag;
  * raw sum (count_abs_lt4_raw / cb_raw_decode): k additions of 0x3F800000 (the bits of 1.0f) into a 32-bit word are decoded by
    ((raw >> 23) * 383) & 511 for every k <= 511.
"""
import numpy as np


def test_unsigned_bit_compare_is_the_interval_test():
    rng = np.random.default_rng(1)
    for delta in (0.5, 2.0, 1e-3, 123.456):
        two_delta = np.float32(2.0) * np.float32(delta)
        k = two_delta.view(np.uint32)
        special = np.array([0.0, -0.0, two_delta, np.nextafter(two_delta, np.float32(0)), np.nextafter(two_delta, np.float32(np.inf)),
                            np.inf, -np.inf, np.nan, 1e-45, -1e-45, 3.4e38, -3.4e38], dtype=np.float32)
        s = np.concatenate([special, rng.normal(delta, 3 * delta, 200_000).astype(np.float32),
                            rng.uniform(-1000, 1000, 100_000).astype(np.float32)])
        bits = s.view(np.uint32)
        inlier_bits = bits < k
        inlier_real = (s >= 0) & (s < two_delta) & ~np.signbit(s)      # -0.0 (s = -delta exactly) is an outlier, as |s| < delta says
        assert np.array_equal(inlier_bits, inlier_real)
        carry = (bits.astype(np.uint64) + np.uint64(2**32 - int(k))) >> np.uint64(32)
        assert np.array_equal(carry.astype(bool), ~inlier_bits)
        # the counter is the high word of a 64-bit accumulator; two steps of the chain
        acc = np.uint64(7) << np.uint64(32)
        for b in bits[:1000]:
            acc = ((acc >> np.uint64(32)) << np.uint64(32) | np.uint64(b)) + np.uint64(2**32 - int(k))
        assert int(acc >> np.uint64(32)) == 7 + int((~inlier_bits[:1000]).sum())


def test_shifted_threshold_equals_abs_threshold_in_exact_arithmetic():
    """|s| < delta  <=>  0 < s + delta < 2 delta; the kernels evaluate the right-hand side with +delta folded into the hoisted
    constant (one rounding at the end of the FFMA chain, as for s itself).  In float64 the two are the same set away from the
    rounding band the fp32 tests allow."""
    rng = np.random.default_rng(2)
    delta = 0.5
    s = rng.normal(0, 1, 1_000_000)
    a = np.abs(s) < delta
    sp = (s + delta).astype(np.float32)
    b = sp.view(np.uint32) < np.float32(2 * delta).view(np.uint32)
    differ = a != b
    assert differ.sum() <= 4 and np.all(np.abs(np.abs(s[differ]) - delta) < 1e-6)


def test_raw_sum_decode():
    one = np.uint32(0x3F800000)
    assert (127 * 383) % 512 == 1
    for k in range(512):
        raw = np.uint32((int(one) * k) & 0xFFFFFFFF)
        assert ((int(raw) >> 23) * 383) & 511 == k
