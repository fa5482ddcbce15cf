"""CPU restatement of the device-RNG resample mode (gravity2_b200/csrc/philox.cu).  TEST
INFRASTRUCTURE ONLY.  Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as
easy as 1, 2, 3", SC'11): counter = (draw >> 2, replicate, 0, 0), key = (seed lo, seed hi),
index = (u32 * P) >> 32.  This is NOT the reference's stream (the reference uses numpy's global
MT19937); it pins the device generator bit for bit.  Known answer from the
Random123 distribution (kat_vectors): philox4x32-10, ctr = key = 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(c, k):
    c0, c1, c2, c3 = c
    k0, k1 = k
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def philox_indices(seed: int, n_boot: int, p: int) -> np.ndarray:
    out = np.empty((n_boot, p), dtype=np.int64)
    key = (seed & MASK, (seed >> 32) & MASK)
    for b in range(n_boot):
        for q in range((p + 3) // 4):
            r = philox4x32_10((q, b, 0, 0), key)
            for i in range(4):
                k = q * 4 + i
                if k < p:
                    out[b, k] = (r[i] * p) >> 32
    return out
