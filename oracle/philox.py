"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the device RNG (pfb_common.cuh): Philox4x32-10
keyed by (seed, offset, index, slot), and the 53-bit uniform built from it.  The reference has no
counterpart (its draws come from numpy's global / per-call generators); this pins the CUDA generator
to the published Philox algorithm (Salmon et al., SC'11) and lets tests regenerate the in-kernel draws."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: (n, 4) uint32, key: (2,) uint32 -> (n, 4) uint32"""
    c = [ctr[:, i].astype(np.uint64) for i in range(4)]
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [(hi1 ^ c[1] ^ np.uint64(k0)) & MASK, lo1, (hi0 ^ c[3] ^ np.uint64(k1)) & MASK, lo0]
        with np.errstate(over="ignore"):
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return np.stack([x.astype(np.uint32) for x in c], axis=1)


def rng_block(seed, offset, index, slot):
    index = np.asarray(index, dtype=np.uint64)
    ctr = np.stack([(index & MASK).astype(np.uint32), (index >> np.uint64(32)).astype(np.uint32),
                    np.full(index.shape, slot, dtype=np.uint32), np.full(index.shape, offset & 0xFFFFFFFF, dtype=np.uint32)], axis=1)
    key = (np.uint32(seed & 0xFFFFFFFF), np.uint32(((seed >> 32) ^ (offset >> 32)) & 0xFFFFFFFF))
    return philox4x32_10(ctr, key)


def u53(a, b):
    return (((a.astype(np.uint64) << np.uint64(32)) | b.astype(np.uint64)) >> np.uint64(11)).astype(np.float64) * 2.0**-53


def resample_uniforms(seed, offset, n):
    """the u_j pfb_resample_rng uses (slot 7)"""
    r = rng_block(seed, offset, np.arange(n, dtype=np.uint64), 7)
    return u53(r[:, 0], r[:, 1])


def systematic_u0(seed, offset):
    r = rng_block(seed, offset, np.zeros(1, dtype=np.uint64), 7)
    return float(u53(r[:, 0], r[:, 1])[0])


def normals(seed, offset, n, D):
    """the standard normals of PFB_RNG_NORMAL, (n, D)"""
    out = np.empty((n, D))
    idx = np.arange(n, dtype=np.uint64)
    for c in range(0, D, 2):
        r = rng_block(seed, offset, idx, c >> 1)
        u1 = 1.0 - u53(r[:, 0], r[:, 1])
        u2 = u53(r[:, 2], r[:, 3])
        rad = np.sqrt(-2.0 * np.log(u1))
        out[:, c] = rad * np.cos(2 * np.pi * u2)
        if c + 1 < D:
            out[:, c + 1] = rad * np.sin(2 * np.pi * u2)
    return out
