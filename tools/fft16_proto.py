"""numpy check of the radix-16 (two fused radix-4 levels per thread) mapping used by the N=1024 FFT kernels:
same butterflies as five in-place radix-4 DIF stages, executed as R16(256,64) -> R16(16,4) -> R4(1)."""
import numpy as np
from fft_proto import digit_reverse4

N = 1024


def tw(span, k, r, sgn=-1.0):
    return np.exp(sgn * 2j * np.pi * (r * k) / (4 * span))


def dif4(a, b, c, d, span, k):
    t0, t1, t2, t3 = a + c, a - c, b + d, (b - d) * (-1j)
    return t0 + t2, (t1 + t3) * tw(span, k, 1), (t0 - t2) * tw(span, k, 2), (t1 - t3) * tw(span, k, 3)


def dit4(a, b, c, d, span, k):
    b, c, d = b * tw(span, k, 1, 1.0), c * tw(span, k, 2, 1.0), d * tw(span, k, 3, 1.0)
    t0, t1, t2, t3 = a + c, a - c, b + d, (b - d) * (1j)
    return t0 + t2, t1 + t3, t0 - t2, t1 - t3


def r16_dif(v, spanA, kA, spanB, kB):
    for m in range(4):
        v[m], v[m + 4], v[m + 8], v[m + 12] = dif4(v[m], v[m + 4], v[m + 8], v[m + 12], spanA, kA(m))
    for blk in range(4):
        v[4 * blk], v[4 * blk + 1], v[4 * blk + 2], v[4 * blk + 3] = dif4(*v[4 * blk:4 * blk + 4], spanB, kB)


def r16_dit(v, spanB, kB, spanA, kA):
    for blk in range(4):
        v[4 * blk], v[4 * blk + 1], v[4 * blk + 2], v[4 * blk + 3] = dit4(*v[4 * blk:4 * blk + 4], spanB, kB)
    for m in range(4):
        v[m], v[m + 4], v[m + 8], v[m + 12] = dit4(v[m], v[m + 4], v[m + 8], v[m + 12], spanA, kA(m))


def forward(x):
    x = x.astype(np.complex128).copy()
    for t in range(64):                       # R16(256, 64): thread t holds idx t + 64 q
        idx = [t + 64 * q for q in range(16)]
        v = [x[i] for i in idx]
        r16_dif(v, 256, lambda m: t + 64 * m, 64, t)
        for i, val in zip(idx, v):
            x[i] = val
    for u in range(64):                       # R16(16, 4): thread u holds idx 64 B + k3 + 4 q
        B, k3 = u // 4, u % 4
        idx = [64 * B + k3 + 4 * q for q in range(16)]
        v = [x[i] for i in idx]
        r16_dif(v, 16, lambda m: k3 + 4 * m, 4, k3)
        for i, val in zip(idx, v):
            x[i] = val
    for w in range(64):                       # R4(1): thread w holds idx 16 w + e
        for m in range(4):
            i0 = 16 * w + 4 * m
            x[i0], x[i0 + 1], x[i0 + 2], x[i0 + 3] = dif4(x[i0], x[i0 + 1], x[i0 + 2], x[i0 + 3], 1, 0)
    return x


def inverse(x):
    x = x.astype(np.complex128).copy()
    for w in range(64):
        for m in range(4):
            i0 = 16 * w + 4 * m
            x[i0], x[i0 + 1], x[i0 + 2], x[i0 + 3] = dit4(x[i0], x[i0 + 1], x[i0 + 2], x[i0 + 3], 1, 0)
    for u in range(64):
        B, k3 = u // 4, u % 4
        idx = [64 * B + k3 + 4 * q for q in range(16)]
        v = [x[i] for i in idx]
        r16_dit(v, 4, k3, 16, lambda m: k3 + 4 * m)
        for i, val in zip(idx, v):
            x[i] = val
    for t in range(64):
        idx = [t + 64 * q for q in range(16)]
        v = [x[i] for i in idx]
        r16_dit(v, 64, t, 256, lambda m: t + 64 * m)
        for i, val in zip(idx, v):
            x[i] = val
    return x


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    X = np.fft.fft(x)
    rev = digit_reverse4(N)
    Y = forward(x)
    print("radix-16 forward vs fft (digit reversed):", np.abs(Y - X[rev]).max())
    print("roundtrip:", np.abs(inverse(Y) / N - x).max())
    # bank check of the two shared-memory exchanges (float2 elements, 16 double-banks, half-warp = 16 lanes)
    def swzA(i): return i ^ (4 * ((i >> 6) & 3))
    def padB(i): return i + (i >> 4)
    worst = 0
    for q in range(16):
        for half in range(4):
            lanes = range(16 * half, 16 * half + 16)
            wr = {swzA(t + 64 * q) % 16 for t in lanes}                      # exchange 1 write (thread t)
            rd = {swzA(64 * (u // 4) + (u % 4) + 4 * q) % 16 for u in lanes}  # exchange 1 read (thread u)
            w2 = {padB(64 * (u // 4) + (u % 4) + 4 * q) % 16 for u in lanes}  # exchange 2 write
            r2 = {padB(16 * w + q) % 16 for w in lanes}                        # exchange 2 read
            worst = max(worst, 16 - len(wr), 16 - len(rd), 16 - len(w2), 16 - len(r2))
    print("bank conflicts (0 = none):", worst)
