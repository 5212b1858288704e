"""Oracle (test infrastructure): the counter-based RNG used by the device samplers.

There is no reference counterpart (the reference uses torch's global Mersenne/Philox generators through
torch.randn / torch.rand, eeyore/samplers/hmc.py:134,148, mala.py:66); this file pins the *builder-defined*
stream layout of ``eeyore_b200/csrc/philox.cuh`` so tests can check device draws bit-for-bit at the integer
level and to rounding at the normal level.

Stream layout (Philox4x32-10, Salmon et al. 2011):
  key     = (seed & 0xffffffff, seed >> 32)
  counter = (block j, iteration t, chain id c, kind)     kind 0 = normals, 1 = accept uniform
  normals : block j yields 4 words (w0..w3).
            fp64: u1 = ((w0<<32|w1)>>11 + 0.5) 2^-53, u2 likewise from (w2,w3) -> Box-Muller pair
                  z[2j] = r cos(2 pi u2), z[2j+1] = r sin(2 pi u2), r = sqrt(-2 log u1)
            fp32: (w0,w1) -> pair z[4j], z[4j+1]; (w2,w3) -> pair z[4j+2], z[4j+3], u = ((w>>8)+0.5) 2^-24
  uniform : block 0 of kind 1; fp64 from (w0,w1), fp32 from w0 (same mappings as above).
"""
from __future__ import annotations

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: 4 arrays of uint32 (broadcastable); key: (k0, k1) ints.  Returns 4 uint32 arrays."""
    c = [np.asarray(a, dtype=np.uint64) for a in np.broadcast_arrays(*ctr)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c[0]
        p1 = _M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return [a.astype(np.uint32) for a in c]


def _u53(hi, lo):
    v = (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
    return ((v >> np.uint64(11)).astype(np.float64) + 0.5) * 2.0 ** -53


def _u24(w):
    """fp32 uniform in (0,1): (k + 1/2) 2^-23 with the top 23 bits k of the word (exact in fp32)."""
    return ((w >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)


def _box_muller(u1, u2):
    dt = u1.dtype
    r = np.sqrt(dt.type(-2) * np.log(u1))
    a = dt.type(2 * np.pi) * u2
    return r * np.cos(a), r * np.sin(a)


def chain_normals(seed, chains, t, P, dtype=np.float64):
    """Standard normals z [len(chains), P] of iteration t."""
    chains = np.asarray(chains, dtype=np.uint32)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    dtype = np.dtype(dtype)
    per = 2 if dtype == np.float64 else 4
    nb = (P + per - 1) // per
    j = np.arange(nb, dtype=np.uint32)[None, :]
    w = philox4x32_10((j, np.uint32(t), chains[:, None], np.uint32(0)), key)
    out = np.empty((chains.shape[0], nb * per), dtype=dtype)
    if dtype == np.float64:
        a, b = _box_muller(_u53(w[0], w[1]), _u53(w[2], w[3]))
        out[:, 0::2], out[:, 1::2] = a, b
    else:
        a, b = _box_muller(_u24(w[0]), _u24(w[1]))
        c, d = _box_muller(_u24(w[2]), _u24(w[3]))
        out[:, 0::4], out[:, 1::4], out[:, 2::4], out[:, 3::4] = a, b, c, d
    return out[:, :P]


def chain_uniforms(seed, chains, t, dtype=np.float64):
    """Accept-test uniforms u [len(chains)] of iteration t, in (0,1)."""
    chains = np.asarray(chains, dtype=np.uint32)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    w = philox4x32_10((np.uint32(0), np.uint32(t), chains, np.uint32(1)), key)
    if np.dtype(dtype) == np.float64:
        return _u53(w[0], w[1])
    return _u24(w[0])
