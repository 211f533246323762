"""TEST INFRASTRUCTURE ONLY.  numpy twin of the engine's dropout RNG (visdial_b200/csrc/common.cuh):
Philox4x32-10, counter = (q_lo, q_hi, site, iteration), key = (seed_lo, seed_hi), q = element index
// 4, word = element index % 4; an element is KEPT iff word >= p * 2^32."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = int(k0), int(k1)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & MASK, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def keep_mask(seed: int, iteration: int, site: int, n: int, p: float) -> np.ndarray:
    """0/1 float32 mask of n elements (row-major linear index of the reference tensor)."""
    nq = (n + 3) // 4
    q = np.arange(nq, dtype=np.uint64)
    out = philox4x32_10(q & MASK, q >> np.uint64(32), np.full(nq, site, np.uint64),
                        np.full(nq, iteration & 0xFFFFFFFF, np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    words = np.stack(out, 1).reshape(-1)[:n]
    thresh = np.uint64(min(int(p * 4294967296.0), 4294967295))
    return (words >= thresh).astype(np.float32)


def make_mask_fn(seed: int, iteration: int, p_of_site=None):
    """mask_fn(site, shape) for oracle.Ctx; p is 0.5 for every site unless p_of_site says otherwise."""
    import torch

    def fn(site, shape):
        n = int(np.prod(shape))
        p = 0.5 if p_of_site is None else p_of_site.get(site, 0.5)
        return torch.from_numpy(keep_mask(seed, iteration, site, n, p)).view(*shape)
    return fn
