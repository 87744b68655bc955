"""The density insert divides by the grid extent as a multiplication by its reciprocal plus two Markstein corrections
(adhoc-queries-pointclouds_b200/csrc/grid_math.cuh, div_by_reciprocal).  The reference's `/` (grid_sampling.rs:51-57) is
the IEEE quotient, so the sequence has to return RN(n / d) bit for bit.  Here the five operations are evaluated in exact
rational arithmetic with one rounding each — what the GPU's DMUL / DFMA do — on random and adversarial operands
(quotients next to integers and to half-way points, divisors just below powers of two, the whole exponent range the
device admits) and compared with Python's correctly rounded float division.  The device code itself is compared with the
oracle by every density parity test."""
import math
import random
import struct
from fractions import Fraction as F


def fma(a, b, c):
    return float(F(a) * F(b) + F(c))  # one rounding (Fraction -> float is correctly rounded)


def div_by_reciprocal(n, d, y):
    q0 = n * y
    r0 = fma(-d, q0, n)
    q1 = fma(r0, y, q0)
    r1 = fma(-d, q1, n)
    return fma(r1, y, q1)


def step(x, k):
    (b,) = struct.unpack("<q", struct.pack("<d", x))
    return struct.unpack("<d", struct.pack("<q", b + k))[0]


def test_five_operation_division_is_the_ieee_quotient():
    rng = random.Random(20261018)
    checked = 0
    for t in range(60_000):
        mode = t % 5
        if mode == 0:  # what the path sees: extents of metres to kilometres, numerators up to extent * cells
            d = rng.uniform(1e-3, 1e5)
            n = rng.uniform(0.0, 1e9)
        elif mode == 1:  # the admitted exponent range, both signs
            d = math.ldexp(rng.uniform(1, 2), rng.randint(-500, 499))
            n = math.ldexp(rng.uniform(1, 2), rng.randint(-500, 499)) * rng.choice([1, -1])
        elif mode == 2:  # quotients within a few ulps of an integer (a point on a cell face)
            d = rng.uniform(1.0, 100.0)
            n = step(rng.randint(1, 5000) * d, rng.randint(-3, 3))
        elif mode == 3:  # divisors just below a power of two, quotients at integers and half-way points
            d = step(2.0 ** rng.randint(-3, 8), -rng.randint(1, 4))
            n = step((rng.randint(1, 5000) + rng.choice([0, 0.5])) * d, rng.randint(-3, 3))
        else:  # significands of all ones / one
            d = step(2.0 ** rng.randint(-10, 10), rng.choice([-1, 0, 1]))
            n = step(2.0 ** rng.randint(-10, 30), rng.choice([-1, 0, 1])) * rng.choice([1, 3, 7])
        if n == 0.0 or not (2.0 ** -500 <= abs(n) < 2.0 ** 500):
            continue
        assert div_by_reciprocal(n, d, 1.0 / d) == n / d, (n, d)
        checked += 1
    assert checked > 55_000


def test_zero_numerators_give_zero():
    for d in (0.1, 3.0, 517.0 * 0.1):
        assert div_by_reciprocal(0.0, d, 1.0 / d) == 0.0
        assert math.copysign(1.0, div_by_reciprocal(-0.0, d, 1.0 / d)) in (1.0, -1.0)  # +-0: cell 0 either way
