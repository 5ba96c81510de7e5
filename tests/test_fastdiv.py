"""The prepared-divisor arithmetic of the keyswitch index decoding (hexl-fpga_b200/csrc/launch.h, FastDiv):
x / d == hi64(x * m) with m = floor((2^64 - 1) / d) + 1 for every 32-bit x and every divisor d >= 2 -- checked here on
the formula itself (edge values of x around every multiple boundary, the largest indices, powers of two, the divisors
the stage jobs really use); the GPU keyswitch tests check the code that uses it bit for bit."""
import random


def fdiv(x, d):
    if d <= 1:
        return x
    m = ((1 << 64) - 1) // d + 1
    return (x * m) >> 64


def test_fastdiv_exact_on_edges_and_random():
    rng = random.Random(7)
    divisors = {2, 3, 5, 6, 7, 8, 36, 42, 49, 63, 64, 1024, 4096 * 6, 16384 * 63, (1 << 31) - 1, (1 << 32) - 1}
    divisors |= {rng.randrange(2, 1 << 32) for _ in range(200)}
    for d in sorted(divisors):
        xs = {0, 1, d - 1, d, d + 1, (1 << 32) - 1, (1 << 32) - d, ((1 << 32) - 1) // d * d, ((1 << 32) - 1) // d * d - 1}
        xs |= {k * d + o for k in (1, 2, 1000, ((1 << 32) - 1) // d) for o in (-1, 0, 1)}
        xs |= {rng.randrange(0, 1 << 32) for _ in range(300)}
        for x in xs:
            if 0 <= x < (1 << 32):
                assert fdiv(x, d) == x // d, (x, d)


def test_stage_index_decoding_matches_plain_division():
    """order / decode of stage S2 (keyswitch_kernels.cu, JobNtt1) with the prepared divisors against // and %"""
    for D in (1, 2, 3, 6, 7):
        for B in (1, 5, 444, 1024):
            lo = B * (D - 1)
            seen = set()
            for i in range(B * D * D):
                if i < D * lo:
                    r = fdiv(i, lo)
                    rem = i - r * lo
                    b = fdiv(rem, D - 1)
                    y = r * (D - 1) + (rem - b * (D - 1))
                    assert (r, b) == (i // lo, (i % lo) // (D - 1))
                else:
                    rem = i - D * lo
                    b = fdiv(rem, D)
                    y = D * (D - 1) + (rem - b * D)
                item = b * D * D + y
                assert item not in seen
                seen.add(item)
                bb = fdiv(item, D * D)
                yy = item - bb * D * D
                assert (bb, yy) == (item // (D * D), item % (D * D))
            assert len(seen) == B * D * D        # the modulus-major walk is a permutation of the items
