'''
CPU study for a round-2 idea: replace the 3xTF32 split of the policy / backup GEMM (3 MMAs at the TF32 rate) by a 2-way FP16
split with power-of-two block scaling (3 MMAs at the FP16 rate = half the tensor time).

    x = s * a (row scale s = 2^k so that max = 2^14),  a1 = fp16(x),  a2 = fp16((x - a1) * 2^11)
    y = t * b (column scale likewise),                 b1 = fp16(y),  b2 = fp16((y - b1) * 2^11)
    a.b = (a1.b1 + 2^-11 (a1.b2 + a2.b1)) / (s t)      (a2.b2 dropped: 2^-22)

Emulated in NumPy (fp16 rounding incl. subnormals, exact products, float64 accumulation) on beliefs produced by the oracle's
belief filter on the T1 model and on alpha vectors with the value range of a trained value function.  Prints the worst relative
error against float64, next to the same figure for the 3xTF32 scheme.  No GPU needed.
'''
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pow2_scale(m, target=14):
    with np.errstate(divide='ignore'):
        k = target - np.ceil(np.log2(np.where(m > 0, m, 1.0)))
    return np.exp2(k)


def split16(x):
    x1 = x.astype(np.float16)
    x2 = ((x - x1.astype(np.float64)) * 2.0 ** 11).astype(np.float16)
    return x1.astype(np.float64), x2.astype(np.float64)


def tf32(x):
    u = np.asarray(x, dtype=np.float32).view(np.uint32)
    return ((u + 0x1000) & 0xFFFFE000).view(np.float32).astype(np.float64)


def main():
    rng = np.random.default_rng(0)
    S, N, G = 30793, 256, 75
    # beliefs with a realistic dynamic range: products of observation likelihoods over a few steps, some peaked, some flat
    B = np.empty((N, S))
    for i in range(N):
        logp = np.zeros(S)
        for _ in range(int(rng.integers(0, 40))):
            c = rng.integers(0, S)
            d = np.abs(np.arange(S) - c) / S
            logp += np.where(rng.random() < 0.3, -40.0 * d, np.log1p(-0.9 * np.exp(-60.0 * d)))
        p = np.exp(logp - logp.max())
        B[i] = p / p.sum()
    B = B.astype(np.float32).astype(np.float64)
    A = (100.0 * rng.random((G, S)) ** 3 * (rng.random((G, S)) < 0.7)).astype(np.float32).astype(np.float64)   # values in [0, 1/(1-gamma)]
    exact = B @ A.T
    norm = np.abs(B) @ np.abs(A).T

    s = pow2_scale(B.max(axis=1))[:, None]
    t = pow2_scale(A.max(axis=1))[:, None]
    a1, a2 = split16(B * s)
    b1, b2 = split16(A * t)
    approx = (a1 @ b1.T + 2.0 ** -11 * (a1 @ b2.T + a2 @ b1.T)) / (s * t.T)
    e16 = np.max(np.abs(approx - exact) / norm)

    ah = tf32(B); al = tf32(B - ah)
    bh = tf32(A); bl = tf32(A - bh)
    e32 = np.max(np.abs(ah @ bh.T + al @ bh.T + ah @ bl.T - exact) / norm)
    flushed = float(np.mean((a1 == 0) & (B > 0)))
    print(f'beliefs: max/min positive ratio up to {np.max(B.max(1) / np.where(B > 0, B, 1).min(1)):.1e}; entries flushed to zero by fp16: {100 * flushed:.2f}%')
    print(f'worst |error| / sum|a||b|:  2-way fp16 split (3 MMAs, fp16 rate) {e16:.2e}   3xTF32 (3 MMAs, tf32 rate) {e32:.2e}')
    print(f'argmax agreement with float64: fp16 split {np.mean(np.argmax(approx, 1) == np.argmax(exact, 1)):.4f}')


if __name__ == '__main__':
    main()
