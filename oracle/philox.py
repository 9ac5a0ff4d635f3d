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


VM_STRIPS = 4096


def vm_strip_table(kappa, K=VM_STRIPS):
    """numpy restatement of pfb_vm_strip_table (pfb_predict_loglik.cu): K strips [a_k, a_k + w_k) of [0, pi] whose
    hats g(a_k) w_k, g(t) = exp(-2 kappa sin^2(t/2)), share the area A -- the smallest A whose K-th strip reaches pi"""

    def end(A, table=None):
        a = 0.0
        for k in range(K):
            with np.errstate(divide="ignore"):  # g underflows far in the tail of a sharp density: w = inf, clipped
                w = A / np.exp(-2.0 * kappa * np.sin(0.5 * a) ** 2)
            if table is not None:
                if not a < np.pi:
                    a, w = np.pi, 0.0
                elif not (a + w < np.pi) or k == K - 1:
                    w = np.pi - a
                table[k] = (a, w)
            elif not (a + w < np.pi):
                return 4.0
            a += w
        return a

    lo, hi = 0.0, float(np.pi)
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if mid <= lo or mid >= hi:
            break
        if end(mid) >= np.pi:
            hi = mid
        else:
            lo = mid
    table = np.zeros((K, 2))
    end(hi, table)
    return table, hi


def _stream_bits(words, start, length):
    """bits [start, start + length) of the little-endian bit stream made of the uint32 columns of `words` (length <= 32)"""
    i, o = start >> 5, start & 31
    lo = words[:, i].astype(np.uint64)
    hi = words[:, i + 1].astype(np.uint64) if i + 1 < words.shape[1] else np.zeros_like(lo)
    return ((lo | (hi << np.uint64(32))) >> np.uint64(o)) & np.uint64((1 << length) - 1)


def _vm_attempt(words, s0, kap, tables):
    """one attempt per row from the 85 bits at s0: (theta >= 0, sign bit, accepted) -- pfb_common.cuh vonmises_attempt"""
    v32 = _stream_bits(words, s0, 32)
    strip = _stream_bits(words, s0 + 32, 12)
    sign = _stream_bits(words, s0 + 44, 1).astype(bool)
    u = ((_stream_bits(words, s0 + 45, 20) << np.uint64(20)) | _stream_bits(words, s0 + 65, 20)).astype(np.float64) * 2.0**-40
    if kap < 1e-8:
        return np.pi * u, sign, np.ones(u.shape, dtype=bool)
    table, A = tables[kap]
    t = table[strip.astype(np.int64)]
    th = t[:, 0] + u * t[:, 1]
    v = (v32.astype(np.float64) + 0.5) * 2.0**-32
    return th, sign, v * A < np.exp(-2.0 * kap * np.sin(0.5 * th) ** 2) * t[:, 1]


def vonmises(seed, offset, n, kappas, tables):
    """the zero-mean von Mises draws of PFB_RNG_VONMISES, (n, D): vertical-strip rejection.  The first attempts of
    all dimensions are cut from one bit stream of ceil(85 D / 128) Philox blocks (slots 16, 17, ...), 85 bits each;
    a rejected dimension c retries with a block of its own (slot 48 + 8 attempt + c).  tables: {kappa: (table, A)}.
    The acceptance test is the fp64 one (the kernel decides in fp32 only where fp32 cannot get it wrong)."""
    D = len(kappas)
    out = np.empty((n, D))
    idx = np.arange(n, dtype=np.uint64)
    nb = (85 * D + 127) // 128
    words = np.concatenate([rng_block(seed, offset, idx, 16 + b) for b in range(nb)], axis=1)
    for c, kap in enumerate(kappas):
        th, sign, acc = _vm_attempt(words, 85 * c, kap, tables)
        out[:, c] = np.where(sign, th, -th)
        pending = np.nonzero(~acc)[0]
        attempt = 1
        while pending.size:
            w = rng_block(seed, offset, idx[pending], 48 + 8 * attempt + c)
            th, sign, acc = _vm_attempt(w, 0, kap, tables)
            out[pending[acc], c] = np.where(sign[acc], th[acc], -th[acc])
            pending = pending[~acc]
            attempt += 1
    return out
