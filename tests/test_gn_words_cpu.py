"""Host-side model of the statistics words used by the convolutions that normalise their own output
(rule_guided_music_b200/csrc/gemm_tc.cuh: gn_publish / gn_ready / gn_affine): per (image, GroupNorm group) two 64-bit
words, bits 63..8 = the sum (sum of squares) as a 36.20 fixed-point integer, bits 7..0 = number of contributions.  One
atomic add delivers a warp's partial AND its arrival.  The properties the device code relies on are checked here with
numpy int64 arithmetic: order independence (bit-identical totals), exact counts, negative sums, and the range the
overflow flag (|partial| <= 1e8) guarantees for up to 255 contributions."""
import numpy as np

FIX = float(1 << 20)
CNT_BITS = 8


def word(partial):
    """What one warp adds for a partial sum (gn_publish)."""
    q = np.int64(np.rint(np.float64(np.float32(partial)) * FIX))
    return ((q << np.int64(CNT_BITS)) + np.int64(1)).astype(np.uint64)  # two's complement: the low byte is the count


def unpack(w):
    """(count, sum) of an accumulator word (gn_ready / gn_affine): arithmetic shift keeps the sign."""
    w = np.uint64(w)
    count = int(w & np.uint64(255))
    total = float(np.int64(w.astype(np.int64)) >> np.int64(CNT_BITS)) / FIX
    return count, total


def accumulate(partials, order):
    acc = np.uint64(0)
    with np.errstate(over="ignore"):
        for i in order:
            acc = np.uint64(acc + word(partials[i]))  # wraps modulo 2^64 like the device's 64-bit atomic add
    return acc


def test_totals_do_not_depend_on_arrival_order_and_count_is_exact():
    rng = np.random.default_rng(0)
    partials = (rng.standard_normal(128) * 300.0).astype(np.float32)  # positive and negative partial sums
    a = accumulate(partials, range(128))
    b = accumulate(partials, rng.permutation(128))
    c = accumulate(partials, reversed(range(128)))
    assert a == b == c, "integer addition is associative: any arrival order gives the same word"
    count, total = unpack(a)
    assert count == 128
    exact = float(np.sum(np.rint(partials.astype(np.float64) * FIX)) / FIX)
    assert total == exact
    assert abs(total - float(partials.astype(np.float64).sum())) <= 128 * 0.5 / FIX


def test_negative_sum_and_partial_counts():
    partials = np.array([-1.5, -2.25, 0.125], dtype=np.float32)
    for k in range(1, 4):
        count, total = unpack(accumulate(partials, range(k)))
        assert count == k, "a word is complete only when its count reaches the image's row tiles x 2"
        assert total == float(partials[:k].astype(np.float64).sum())


def test_range_guaranteed_by_the_overflow_flag():
    """gn_publish flags partials above 1e8; 255 contributions of that size must still fit bits 63..8 with the sign."""
    limit = 1.0e8
    worst = 255 * np.rint(limit * FIX)
    assert worst < 2.0 ** 55, "36.20 fixed point in 56 bits: |sum| < 2^35"
    partials = np.full(255, limit, dtype=np.float32)
    count, total = unpack(accumulate(partials, range(255)))
    assert count == 255
    assert total == 255 * float(np.float32(limit))
    count, total = unpack(accumulate(-partials, range(255)))
    assert count == 255 and total == -255 * float(np.float32(limit))


def test_affine_from_totals_matches_group_norm():
    """gn_affine: mean = s / n, var = q / n - mean^2, y = (x - mean) rstd gamma + beta as a x + b."""
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((64, 128, 4)) * 2.0 + 0.7).astype(np.float32)  # 64 row tiles x 128 rows x 4 channels: one group
    s_parts = x.reshape(128, 64 * 4).sum(axis=1)          # 128 contributions
    q_parts = (x.astype(np.float64) ** 2).reshape(128, 64 * 4).sum(axis=1).astype(np.float32)
    cs, s = unpack(accumulate(s_parts, range(128)))
    cq, q = unpack(accumulate(q_parts, range(128)))
    assert cs == cq == 128
    n = x.size
    mean = s / n
    var = q / n - mean * mean
    rstd = 1.0 / np.sqrt(max(var, 0.0) + 1e-6)
    ref_mean, ref_var = float(x.astype(np.float64).mean()), float(x.astype(np.float64).var())
    assert abs(mean - ref_mean) < 1e-5 and abs(var - ref_var) < 1e-4 * ref_var
    y = (x - mean) * rstd
    assert abs(float(y.mean())) < 1e-4 and abs(float(y.std()) - 1.0) < 1e-3
