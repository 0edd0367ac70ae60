"""Pure-Python Philox4x32-10 stream -- TEST INFRASTRUCTURE.

An implementation independent of the C oracle and of the CUDA kernel, used to feed the
unmodified reference (through refharness.StreamShim) the same per-env uniform stream the
product's "philox" RNG mode defines:

    draw j of an env with 64-bit seed S:
        block  = Philox4x32-10(counter = (lo32(j>>1), hi32(j>>1), 0x50434352, 0),
                               key     = (lo32(S), hi32(S)))
        (a, b) = (block[0], block[1]) if j even else (block[2], block[3])
        u      = ((a >> 5) * 2**26 + (b >> 6)) / 2**53          # CPython genrand_res53 form
"""
M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
DOMAIN = 0x50434352
MASK = 0xFFFFFFFF


def philox4x32_10(c, k):
    c0, c1, c2, c3 = c
    k0, k1 = k
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


class PhiloxStream(object):
    def __init__(self, seed):
        self.seed = seed & 0xFFFFFFFFFFFFFFFF
        self.draws = 0
        self._blk = None
        self._blk_idx = -1

    def random(self):
        j = self.draws
        blk = j >> 1
        if blk != self._blk_idx:
            self._blk = philox4x32_10((blk & MASK, (blk >> 32) & MASK, DOMAIN, 0),
                                      (self.seed & MASK, (self.seed >> 32) & MASK))
            self._blk_idx = blk
        a, b = (self._blk[2], self._blk[3]) if (j & 1) else (self._blk[0], self._blk[1])
        self.draws += 1
        return ((a >> 5) * 67108864.0 + (b >> 6)) * (1.0 / 9007199254740992.0)
