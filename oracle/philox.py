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


# ---- "multinomial_sorted" draws (pfb_msort.cu) --------------------------------------------------------------
# n_out i.i.d. uniforms generated leaf by leaf: [0, 1) is cut into K = 2^k equal leaves; the occupancies are
# multinomial(n_out; 1/K, ..), drawn by binomial splitting with p = 1/2 -- a node holding m uniforms hands
# popcount(first m bits of its Philox stream) to its left child -- and the uniforms of leaf b are
# (b 2^(53-k) + r) 2^-53 with fresh (53-k)-bit draws r.  Slot order = leaf order.
MS_LEAF_TARGET = 32
MS_SLOT_TREE = 64
MS_SLOT_LEAF = 33


def msort_levels(n_out):
    k = 0
    while (MS_LEAF_TARGET << k) < n_out:
        k += 1
    return k


def msort_top_levels(k):
    """levels whose nodes are coarse bins (one CTA each): pfb_msort.cu ms_plan"""
    if k <= 4:
        return 0
    return max(k - 9, min(k - 4, 9))


def _popcount32(x):
    x = x.astype(np.uint32)
    x = x - ((x >> np.uint32(1)) & np.uint32(0x55555555))
    x = (x & np.uint32(0x33333333)) + ((x >> np.uint32(2)) & np.uint32(0x33333333))
    x = (x + (x >> np.uint32(4))) & np.uint32(0x0F0F0F0F)
    return ((x * np.uint32(0x01010101)) >> np.uint32(24)).astype(np.int64)


def msort_leaf_counts(seed, offset, n_out):
    """occupancy of the 2^k leaves (int64 array), by the popcount tree"""
    k = msort_levels(n_out)
    cnt = np.array([n_out], dtype=np.int64)
    for l in range(k):
        nblk = (cnt + 127) // 128                      # Philox blocks per node
        node = np.repeat(np.arange(cnt.shape[0], dtype=np.uint64), nblk)
        first = np.cumsum(nblk) - nblk
        q = np.arange(int(nblk.sum()), dtype=np.uint64) - np.repeat(first, nblk).astype(np.uint64)
        r = rng_block(seed, offset, (node << np.uint64(32)) | q, MS_SLOT_TREE + l)
        bits = np.minimum(128, np.repeat(cnt, nblk) - q.astype(np.int64) * 128)   # bits used of each block
        pc = np.zeros(q.shape[0], dtype=np.int64)
        for j in range(4):
            b = np.clip(bits - 32 * j, 0, 32)
            mask = np.where(b >= 32, np.uint64(0xFFFFFFFF), (np.uint64(1) << b.astype(np.uint64)) - np.uint64(1))
            pc += _popcount32(r[:, j].astype(np.uint64) & mask)
        left = np.zeros(cnt.shape[0], dtype=np.int64)
        np.add.at(left, node.astype(np.int64), pc)
        nxt = np.empty(2 * cnt.shape[0], dtype=np.int64)
        nxt[0::2] = left
        nxt[1::2] = cnt - left
        cnt = nxt
    return cnt


def msort_uniforms(seed, offset, n_out):
    """the n_out uniforms of pfb_resample_msort in slot order, and the leaf of every slot"""
    k = msort_levels(n_out)
    cnt = msort_leaf_counts(seed, offset, n_out)
    leaf = np.repeat(np.arange(cnt.shape[0], dtype=np.uint64), cnt)
    # in-leaf draws: one Philox block per pair of slots (2 j, 2 j + 1) of a coarse bin (the unit one CTA works on)
    k_top = msort_top_levels(k)
    binc = cnt.reshape(1 << k_top, -1).sum(axis=1)
    bin_id = np.repeat(np.arange(binc.shape[0], dtype=np.uint64), binc)
    s = np.arange(n_out, dtype=np.uint64) - np.repeat(np.cumsum(binc) - binc, binc).astype(np.uint64)
    r = rng_block(seed, offset, (bin_id << np.uint64(32)) | (s >> np.uint64(1)), MS_SLOT_LEAF)
    odd = (s & np.uint64(1)).astype(bool)
    hi = np.where(odd, r[:, 2], r[:, 0]).astype(np.uint64)
    lo = np.where(odd, r[:, 3], r[:, 1]).astype(np.uint64)
    r64 = (hi << np.uint64(32)) | lo
    R = (leaf << np.uint64(53 - k)) | (r64 >> np.uint64(11 + k))
    return R.astype(np.float64) * 2.0**-53, leaf.astype(np.int64)
