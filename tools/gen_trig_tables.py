#!/usr/bin/env python3
"""Generate the double-double constants of csrc/friq_trig.h with mpmath (200 digits)."""
from mpmath import mp, mpf, sin, cos, pi, factorial
import struct

mp.prec = 400


def dbl(x):
    return float(x)


def hexf(x):
    return float(x).hex()


def dd(x):
    hi = mpf(dbl(x)); lo = mpf(dbl(x - hi))
    return hexf(hi), hexf(lo)


def top_bits(x, bits):
    """x truncated to `bits` significant bits (as mpf)."""
    from mpmath import floor, log
    e = int(floor(log(x, 2)))
    q = mpf(2) ** (e - bits + 1)
    return floor(x / q) * q


p = pi / 2
P1 = top_bits(p, 33); r = p - P1
P2 = top_bits(r, 33); r -= P2
P3 = top_bits(r, 33); r -= P3
P4 = mpf(dbl(r))
print("#define FRIQ_PIO2_1 %s" % hexf(P1))
print("#define FRIQ_PIO2_2 %s" % hexf(P2))
print("#define FRIQ_PIO2_3 %s" % hexf(P3))
print("#define FRIQ_PIO2_4 %s" % hexf(P4))
print("#define FRIQ_TWO_OVER_PI %s" % hexf(2 / pi))
for n, name in [(3, "S3"), (5, "S5")]:
    v = (-1) ** ((n - 1) // 2) / factorial(n)
    h, l = dd(v); print("#define FRIQ_%s_HI %s\n#define FRIQ_%s_LO %s" % (name, h, name, l))
for n in (7, 9, 11, 13):
    print("#define FRIQ_S%d %s" % (n, hexf((-1) ** ((n - 1) // 2) / factorial(n))))
h, l = dd(mpf(1) / 24); print("#define FRIQ_C4_HI %s\n#define FRIQ_C4_LO %s" % (h, l))
for n in (6, 8, 10, 12):
    print("#define FRIQ_C%d %s" % (n, hexf((-1) ** (n // 2) / factorial(n))))
print("/* sin(j/32), cos(j/32) as double-double, j = 0..26 : {sin_hi, sin_lo, cos_hi, cos_lo} */")
for j in range(27):
    s = dd(sin(mpf(j) / 32)); c = dd(cos(mpf(j) / 32))
    print("    {%s, %s, %s, %s}," % (s[0], s[1], c[0], c[1]))
