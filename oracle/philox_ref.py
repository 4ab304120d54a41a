"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the counter-based dropout RNG.

The CUDA kernels (csrc/common.cuh: philox4x32_10 / dropout_keep) draw dropout
decisions from Philox4x32-10 (Salmon et al., SC'11) keyed by (seed, stream,
element index).  The reference uses torch's nn.Dropout (clip_model.py:27,65,86,92,238);
its masks cannot be reproduced bit-for-bit by any other RNG, so train-mode
parity is checked by feeding *these* masks to the reference/oracle.

Element index convention (shared with the kernels).  Philox block b = philox(ctr = (b lo, b hi, stream, 0), key = seed).
  p != 0.5 : 16 bits per decision, 8 per Philox call
      block = idx >> 3, e = idx & 7 ; draw = (words[e >> 1] >> (16 * (e & 1))) & 0xffff ; keep = draw >= floor(p * 2**16)
  p == 0.5 : ONE bit per decision, 128 per Philox call (the attention / projection / FFN sites of the encoder)
      block = idx >> 7, e = idx & 127 ; keep = (words[e >> 5] >> (e & 31)) & 1
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. All inputs broadcastable uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint32)
    c1 = np.asarray(c1, dtype=np.uint32)
    c2 = np.asarray(c2, dtype=np.uint32)
    c3 = np.asarray(c3, dtype=np.uint32)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK32).astype(np.uint32)
            n0 = hi1 ^ c1 ^ k0
            n2 = hi0 ^ c3 ^ k1
            c0, c1, c2, c3 = n0, lo1, n2, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def dropout_threshold(p):
    return np.uint32(min(int(np.floor(float(p) * 65536.0)), 0xFFFF))


def keep_mask(n, seed, stream, p):
    """Boolean keep mask for element indices 0..n-1 (flat, C order)."""
    n = int(n)
    onebit = float(np.float32(p)) == 0.5
    per = 128 if onebit else 8
    nb = (n + per - 1) // per
    blk = np.arange(nb, dtype=np.uint64)
    c0 = (blk & MASK32).astype(np.uint32)
    c1 = (blk >> np.uint64(32)).astype(np.uint32)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    w = philox4x32_10(c0, c1, np.uint32(stream), np.uint32(0), seed & 0xFFFFFFFF, seed >> 32)
    words = np.stack(w, axis=1)                                   # (nb, 4)
    if onebit:
        bits = (words[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & np.uint32(1)
        return bits.reshape(-1)[:n].astype(bool)                  # element e -> word e>>5, bit e&31
    lo = words & np.uint32(0xFFFF)
    hi = words >> np.uint32(16)
    draws = np.stack([lo, hi], axis=2).reshape(-1)[:n]            # element e -> word e>>1, half e&1
    return draws >= dropout_threshold(p)


def stream_id(layer, site):
    """stream = layer * 16 + site ; sites: 0 conv, 1 attn-prob, 2 post-proj, 3 ffn-hidden, 4 post-ffn."""
    return int(layer) * 16 + int(site)
