'''
The arithmetic of the 2-way FP16 split behind umma_f16_kernel, emulated in NumPy (fp16 rounding with subnormals, exact products,
float64 accumulation; scripts/fp16_split_study.py): with power-of-two block scales the scheme keeps FP32-class accuracy on
operands whose magnitudes span far more than fp16's range.  CPU only.
'''
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'scripts'))
from fp16_split_study import pow2_scale, split16, tf32   # noqa: E402


def test_block_scaled_fp16_split_is_as_accurate_as_3xtf32():
    rng = np.random.default_rng(3)
    n, S, G = 64, 4000, 24
    B = np.exp(-rng.uniform(0, 80, (n, S))) * (rng.random((n, S)) < 0.8)          # 35 decades inside a row
    B = (B / B.sum(1, keepdims=True) * 10.0 ** rng.uniform(-20, 0, (n, 1))).astype(np.float32).astype(np.float64)   # deferred normalisers
    A = (rng.standard_normal((G, S)) * 10.0 ** rng.integers(-6, 3, (G, 1))).astype(np.float32).astype(np.float64)
    exact, norm = B @ A.T, np.abs(B) @ np.abs(A).T
    s = pow2_scale(B.sum(axis=1))[:, None]                   # the kernel scales a row by its normaliser (entries <= sum)
    t = pow2_scale(np.abs(A).max(axis=1))[:, None]
    a1, a2 = split16(B * s)
    b1, b2 = split16(A * t)
    assert np.all(np.isfinite(a1)) and np.all(np.isfinite(b1)) and np.abs(a1).max() <= 2.0 ** 14 and np.abs(b1).max() <= 2.0 ** 14
    approx = (a1 @ b1.T + 2.0 ** -11 * (a1 @ b2.T + a2 @ b1.T)) / (s * t.T)
    err16 = np.max(np.abs(approx - exact) / norm)
    ah, bh = tf32(B), tf32(A)
    al, bl = tf32(B - ah), tf32(A - bh)
    err32 = np.max(np.abs(ah @ bh.T + al @ bh.T + ah @ bl.T - exact) / norm)
    assert err16 < 5e-7 and err16 < 4 * err32 + 1e-7
    assert np.array_equal(np.argmax(approx, 1), np.argmax(exact, 1))
    # without the block scales the same split loses the rows entirely (every entry is below fp16's smallest subnormal)
    tiny = B[np.argmin(B.sum(1))]
    assert np.all(split16(tiny[None, :])[0] == 0) or tiny.sum() > 1e-7
