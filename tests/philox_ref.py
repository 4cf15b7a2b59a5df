"""numpy restatement of Philox4x32-10 + Box-Muller exactly as csrc/sampler.cu draws it (test helper)."""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c = [int(v) & MASK for v in ctr]
    k0, k1 = int(key[0]) & MASK, int(key[1]) & MASK
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k1) & MASK, p0 & MASK]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c


def philox_normal4(c0, c1, c2, c3, k0, k1):
    c = philox4x32_10((c0, c1, c2, c3), (k0, k1))
    out = np.zeros(4, dtype=np.float64)
    s = 2.0 ** -24
    for p in range(2):
        u1 = ((c[2 * p] >> 8) + 0.5) * s
        u2 = ((c[2 * p + 1] >> 8) + 0.5) * s
        r = np.sqrt(-2.0 * np.log(u1))
        out[2 * p] = r * np.cos(2 * np.pi * u2)
        out[2 * p + 1] = r * np.sin(2 * np.pi * u2)
    return out
