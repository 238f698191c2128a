"""TEST INFRASTRUCTURE ONLY -- pure-Python Philox4x32-10 (Salmon et al., SC'11, public algorithm).

Mirrors the device function ``philox4x32_10`` in ``openmines_b200/csrc/om_rng.cuh`` so that
integer draws (the RandomDispatcher's ``randint``; reference call sites
random_dispatcher.py:16,22,29) are reproducible bit-for-bit on both sides without a tape.

Keying convention: key = (seed_lo, seed_hi) of the instance,
counter = (draw_epoch, entity_id, site_id, 0).
"""
M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF

SITE_DISPATCH = 12


def philox4x32_10(counter, key):
    c0, c1, c2, c3 = [int(x) & MASK for x in counter]
    k0, k1 = [int(x) & MASK for x in key]
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK
        hi1, lo1 = p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def randint_keyed(seed, epoch, entity, n, site=SITE_DISPATCH):
    """Unbiased-enough bounded integer: (x0 * n) >> 32 (Lemire multiply-shift, no rejection)."""
    seed = int(seed)
    x0 = philox4x32_10((epoch, entity, site, 0), (seed & MASK, (seed >> 32) & MASK))[0]
    return (x0 * int(n)) >> 32
