"""NumPy restatement of the per-chain random streams of the CUDA library  --  TEST INFRASTRUCTURE ONLY.

The kernels draw from counter-based Philox4x32-10 streams (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
SC'11) keyed by the chain's 64-bit seed (littlemcmc_b200/csrc/lmc_device.cuh):

    uniform k of transition `it`:      philox(counter = (k, it_lo, it_hi, "UNIF"), key = (seed_lo, seed_hi)) -> words x, y
                                       u = ((y << 32 | x) >> 12) + 0.5) * 2^-52                      in (0, 1)
    normals (2j, 2j+1) of transition:  philox(counter = (j, it_lo, it_hi, "NORM"), key) -> words x, y, z, w
                                       r = sqrt(-2 log u(x, y)),  (r cos(2 pi u(z, w)), r sin(2 pi u(z, w)))

`philox4x32_10` is checked against the Random123 known-answer vectors in tests/test_oracle_golden.py; the GPU tests
require `lmc_rng_fill` to reproduce `uniforms()` bit for bit and `normals()` to 1e-14 (libm log / sincos).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
TAG_UNIFORM, TAG_NORMAL = 0x554E4946, 0x4E4F524D
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: 4 arrays of uint32 (broadcastable), key: 2 uint32 -> 4 uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) & _MASK for x in np.broadcast_arrays(*counter)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & _MASK, p1 >> np.uint64(32), p1 & _MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in c]


def u52(lo, hi):
    m = ((hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)) >> np.uint64(12)
    return (m.astype(np.float64) + 0.5) * 2.0 ** -52


def _it_words(it):
    it = int(it) & 0xFFFFFFFFFFFFFFFF
    return it & 0xFFFFFFFF, it >> 32


def uniforms(seed, it, n):
    """The first n uniforms of transition `it` of the chain keyed by `seed`."""
    lo, hi = _it_words(it)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    r = philox4x32_10((np.arange(n, dtype=np.uint64), lo, hi, TAG_UNIFORM), (seed & 0xFFFFFFFF, seed >> 32))
    return u52(r[0], r[1])


def normals(seed, it, ndim):
    """The ndim standard normals of the momentum draw of transition `it`."""
    lo, hi = _it_words(it)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    pairs = (ndim + 1) // 2
    r = philox4x32_10((np.arange(pairs, dtype=np.uint64), lo, hi, TAG_NORMAL), (seed & 0xFFFFFFFF, seed >> 32))
    rad = np.sqrt(-2.0 * np.log(u52(r[0], r[1])))
    ang = 2.0 * np.pi * u52(r[2], r[3])
    out = np.empty(2 * pairs)
    out[0::2], out[1::2] = rad * np.cos(ang), rad * np.sin(ang)
    return out[:ndim]
