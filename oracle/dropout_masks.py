"""Train-mode dropout masks of the CUDA path, restated on the CPU (numpy).  TEST INFRASTRUCTURE ONLY.

The reference applies nn.Dropout at four kinds of sites (models/qformer.py:107 embeddings, :258 attention
probabilities, :287 attention output dense, :373 FFN output dense) with torch's global RNG; a random mask has no
cross-implementation parity by itself.  The CUDA path therefore DEFINES its masks as a pure function of
(seed, site, element index) through the counter-based generator Philox4x32-10 (Salmon et al., SC'11; the same
generator torch/cuRAND use), so a CPU restatement can reproduce every mask bit and the oracle can run the same
dropped forward/backward.  This file is that restatement; tests pin it against the Random123 known-answer vectors.

Definition (shared with unirec_b200/csrc/dropout.cuh):
  * words = philox4x32_10(counter = (row & 0xffffffff, group, site, row >> 32), key = (seed & 0xffffffff, seed >> 32))
  * the 4 words give 8 16-bit values v[j], j = 0..7: v[j] = (words[j >> 1] >> (16 * (j & 1))) & 0xffff
  * element kept iff v[j] >= thr16, thr16 = round(p * 65536); kept elements are scaled by 65536 / (65536 - thr16)
  * hidden-state sites ([rows, H] tensors): row = row index, group = col >> 3, j = col & 7
  * attention-probability sites ([B, heads, nq, nk]): row = (b * heads + h) * nq + q,
    group = (k >> 5) * 4 + ((k & 7) >> 1), j = 2 * ((k & 31) >> 3) + (k & 1)
    (one Philox call = the 8 probabilities one thread of an mma.sync quad owns in a 32-key block)
  * site ids: 0 = embeddings; 1 + 8 * layer + {0: self-attention probs, 1: self-attention output dense,
    2: cross-attention probs, 3: cross-attention output dense, 4: FFN output dense}
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85

SITE_EMBEDDINGS = 0
KIND_SELF_PROBS, KIND_SELF_OUT, KIND_CROSS_PROBS, KIND_CROSS_OUT, KIND_FFN_OUT = range(5)


def site_id(layer: int, kind: int) -> int:
    return 1 + 8 * layer + kind


def threshold16(p: float) -> int:
    return int(round(p * 65536.0))


def keep_scale(thr16: int) -> float:
    return 65536.0 / (65536.0 - thr16)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised Philox4x32-10: counters are uint32 arrays (broadcastable), key two python ints."""
    c0, c1, c2, c3 = np.broadcast_arrays(np.asarray(c0, np.uint32), np.asarray(c1, np.uint32),
                                         np.asarray(c2, np.uint32), np.asarray(c3, np.uint32))
    mask32 = np.uint64(0xFFFFFFFF)
    for r in range(10):
        p0 = PHILOX_M0 * c0.astype(np.uint64)
        p1 = PHILOX_M1 * c2.astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask32).astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask32).astype(np.uint32)
        kk0 = np.uint32((k0 + r * PHILOX_W0) & 0xFFFFFFFF)
        kk1 = np.uint32((k1 + r * PHILOX_W1) & 0xFFFFFFFF)
        c0, c1, c2, c3 = hi1 ^ c1 ^ kk0, lo1, hi0 ^ c3 ^ kk1, lo0
    return c0, c1, c2, c3


def _values16(row, group, j, site: int, seed: int):
    w = philox4x32_10(row & np.uint64(0xFFFFFFFF), group, np.uint32(site), row >> np.uint64(32),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    words = np.stack(w, axis=0)                                      # [4, ...]
    word = np.take_along_axis(words, (j >> 1)[None].astype(np.int64), axis=0)[0]
    return (word >> (16 * (j & 1)).astype(np.uint32)) & np.uint32(0xFFFF)


def keep_mask_rows(seed: int, site: int, rows: int, width: int, thr16: int) -> np.ndarray:
    """bool [rows, width]: True where the element is kept (hidden-state sites).  One Philox call per (row, group of 8
    columns), its 4 words split into the 8 16-bit values of the group (the definition above, evaluated once per call
    instead of once per element)."""
    groups = (width + 7) // 8
    r = np.arange(rows, dtype=np.uint64)[:, None]
    gidx = np.arange(groups, dtype=np.uint32)[None, :]
    r, gidx = np.broadcast_arrays(r, gidx)
    w = philox4x32_10(r & np.uint64(0xFFFFFFFF), gidx, np.uint32(site), r >> np.uint64(32),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    v = np.empty((rows, groups, 8), dtype=np.uint32)
    for i, word in enumerate(w):                       # v[j] = (words[j >> 1] >> (16 * (j & 1))) & 0xffff
        v[:, :, 2 * i] = word & np.uint32(0xFFFF)
        v[:, :, 2 * i + 1] = word >> np.uint32(16)
    return v.reshape(rows, groups * 8)[:, :width] >= np.uint32(thr16)


def keep_mask_attention(seed: int, site: int, batch: int, heads: int, nq: int, nk: int, thr16: int) -> np.ndarray:
    """bool [batch, heads, nq, nk]: True where the probability is kept."""
    row = np.arange(batch * heads * nq, dtype=np.uint64)[:, None]
    k = np.arange(nk, dtype=np.uint32)[None, :]
    row, k = np.broadcast_arrays(row, k)
    group = (k >> 5) * 4 + ((k & 7) >> 1)
    j = 2 * ((k & 31) >> 3) + (k & 1)
    v = _values16(row, group.astype(np.uint32), j.astype(np.uint32), site, seed)
    return (v >= np.uint32(thr16)).reshape(batch, heads, nq, nk)
