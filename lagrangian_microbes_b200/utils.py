"""Small host helpers with the semantics of /root/reference/utils.py (re-implemented)."""
import math
from datetime import datetime, timedelta


def factor(n):
    """Sorted list of the positive divisors of n (utils.py:7-22)."""
    small = [d for d in range(1, math.isqrt(n) + 1) if n % d == 0]
    large = [n // d for d in reversed(small) if d * d != n]
    return small + large


def most_symmetric_integer_factorization(N):
    """The factor pair (a, b), a <= b, a*b == N, closest to a square (utils.py:25-42).

    The reference picks the divisor nearest sqrt(N) with ties to the smaller one; since for a
    divisor pair a < sqrt(N) < b the smaller one is always nearer, that is the largest divisor
    not exceeding sqrt(N).  490000 -> (700, 700).
    """
    N = int(N)
    for a in range(math.isqrt(N), 0, -1):
        if N % a == 0:
            return a, N // a
    raise ValueError("N must be a positive integer")


def closest_hour(ndt):
    """numpy datetime64 -> python datetime rounded to the nearest hour, half up (utils.py:45-54)."""
    seconds = int(ndt.astype("datetime64[s]").astype("int64"))
    base = datetime(1970, 1, 1) + timedelta(seconds=seconds)
    floor = base.replace(minute=0, second=0, microsecond=0)
    return floor + timedelta(hours=1) if base.minute >= 30 else floor


def pretty_time(t):
    """Seconds -> short human string (utils.py:57-67)."""
    for limit, scale, unit in ((1e-6, 1e9, "ns"), (1e-3, 1e6, "μs"), (1.0, 1e3, "ms"), (60.0, 1.0, "s")):
        if t < limit:
            return "{:.3g} {:s}".format(t * scale, unit)
    return "{:.3g} mins".format(t / 60)


def pretty_filesize(num, suffix="B"):
    """Bytes -> binary-prefixed string (utils.py:70-84)."""
    units = ["", "Ki", "Mi", "Gi", "Ti", "Pi", "Ei", "Zi"]
    k = 0
    while abs(num) >= 1024.0 and k < len(units):
        num /= 1024.0
        k += 1
    if k == len(units):
        return "%.1f %s%s" % (num, "Yi", suffix)
    return "%3.1f %s%s" % (num, units[k], suffix)
