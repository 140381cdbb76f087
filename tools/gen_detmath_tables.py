#!/usr/bin/env python
"""Prints the exact Taylor/Maclaurin coefficients (rounded once to double, as C hex-float literals)
and the hi/lo constant splits used by radarays_ros_b200/csrc/rr_detmath.h."""
from decimal import Decimal, getcontext
from fractions import Fraction
from math import factorial


def hexd(fr):
    return float(fr).hex()


def main():
    print("SIN  (-1)^k/(2k+1)!, k=1..10:", [hexd(Fraction((-1) ** k, factorial(2 * k + 1))) for k in range(1, 11)])
    print("COS  (-1)^k/(2k)!,   k=1..11:", [hexd(Fraction((-1) ** k, factorial(2 * k))) for k in range(1, 12)])
    print("ASIN (2k)!/(4^k k!^2 (2k+1)), k=1..29:",
          [hexd(Fraction(factorial(2 * k), 4 ** k * factorial(k) ** 2 * (2 * k + 1))) for k in range(1, 30)])
    print("EXP  1/k!, k=2..15:", [hexd(Fraction(1, factorial(k))) for k in range(2, 16)])
    print("LOG  2/(2k+1), k=0..13:", [hexd(Fraction(2, 2 * k + 1)) for k in range(0, 14)])
    getcontext().prec = 80

    def arctan_inv(n):
        x = Decimal(1) / n
        s = x
        t = x
        k = 1
        while True:
            t = -t / (n * n)
            k += 2
            d = t / k
            if abs(d) < Decimal(10) ** -78:
                return s
            s += d

    pi = 4 * (4 * arctan_inv(5) - arctan_inv(239))
    ln2 = Decimal(2).ln()
    for name, d in (("pi/2", pi / 2), ("pi", pi), ("ln2", ln2)):
        hi = float(d)
        lo = float(d - Decimal(hi))
        print(name, "hi", hi.hex(), "lo", lo.hex())
    print("1/ln2", float(1 / ln2).hex(), "2/pi", float(2 / pi).hex())


if __name__ == "__main__":
    main()
