"""java.util.Random restated (48-bit LCG), for the per-pass seeds.

The reference seeds every pass with ``rand.nextInt()`` drawn from
``new Random(0)`` (OpenClPathTracingRenderer.java:95,106-107).  The algorithm is
the one published in the Java SE API documentation of ``java.util.Random``.
"""

_MULT = 0x5DEECE66D
_MASK = (1 << 48) - 1


class JavaRandom:
    def __init__(self, seed: int = 0):
        self.state = (seed ^ _MULT) & _MASK

    def _next(self, bits: int) -> int:
        self.state = (self.state * _MULT + 0xB) & _MASK
        v = self.state >> (48 - bits)
        if v >= 1 << (bits - 1):          # to signed
            v -= 1 << bits
        return v

    def next_int(self) -> int:
        return self._next(32)


def pass_seeds(n: int, seed: int = 0, skip: int = 0):
    """First ``n`` values of ``new Random(seed).nextInt()`` after ``skip`` draws (signed int32)."""
    r = JavaRandom(seed)
    for _ in range(skip):
        r.next_int()
    return [r.next_int() for _ in range(n)]
