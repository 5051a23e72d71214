"""numpy model of lenia_march_kernel's data flow (pyca_b200/csrc/lenia_march.cu): validates, on the CPU, the index
arithmetic of the 16 x 16 register FFT (thread t, register q <-> sample t + 16 q), the slot layout of the spectra rows
(float4 = two adjacent k_b of one k_a), the tap table the FIR threads read, and the band-local row pairing
(rows y0 + p and y0 + Lp + p ride in one complex transform) against a direct periodic correlation.

    python tools/march_proto.py
"""
import numpy as np

N = 256
R = 15
V = N - 2 * R      # valid outputs per strip
W16 = np.exp(-2j * np.pi / 16)
W256 = np.exp(-2j * np.pi / 256)


def dft16(v, sign=-1):
    """v[..., 16] natural in -> natural out, as two radix-4 stages (the register network of the kernel)."""
    w = np.exp(sign * 2j * np.pi / 16)
    w4 = w ** 4
    B = np.zeros(v.shape, complex)
    for q1 in range(4):
        for ka in range(4):
            acc = 0
            for q2 in range(4):
                acc = acc + v[..., q1 + 4 * q2] * w4 ** (q2 * ka)
            B[..., q1 * 4 + ka] = acc * w ** (q1 * ka)
    out = np.zeros(v.shape, complex)
    for ka in range(4):
        for kb in range(4):
            acc = 0
            for q1 in range(4):
                acc = acc + B[..., q1 * 4 + ka] * w4 ** (q1 * kb)
            out[..., ka + 4 * kb] = acc
    return out


def fwd256(x):
    """x[256] -> regs[ka][kb] = X[ka + 16 kb] (thread ka, register kb)"""
    t = np.arange(16)
    regs = x[t[:, None] + 16 * np.arange(16)[None, :]]          # thread t=n1, reg q=n2
    A = dft16(regs)                                              # reg index ka
    A = A * W256 ** (t[:, None] * np.arange(16)[None, :])        # twiddle W^(n1 ka)
    ex = A.T.copy()                                              # exchange: thread ka, reg n1
    return dft16(ex)                                             # reg kb


def inv256(Y):
    """Y regs[ka][kb] -> y[256] (unnormalised inverse), thread n1 reg n2 -> sample n1 + 16 n2"""
    C = dft16(Y, +1)                                             # thread ka, reg n1
    ka = np.arange(16)
    C = C * np.conj(W256) ** (ka[:, None] * np.arange(16)[None, :])
    ex = C.T.copy()                                              # thread n1, reg ka
    y = dft16(ex, +1)                                            # reg n2
    out = np.zeros(256, complex)
    t = np.arange(16)
    out[t[:, None] + 16 * np.arange(16)[None, :]] = y
    return out


def slot_of(ka, kb):
    """float4 index and lane (0/1) inside the spectra row for frequency ka + 16 kb"""
    j = kb >> 1
    return ka * 8 + j, kb & 1      # (the shared-memory row pads this to 9 * ka + j)


def freq_of_f(f):
    """the two frequencies FIR thread f (float4 index f) owns"""
    ka = f >> 3
    j = f & 7
    return ka + 16 * (2 * j), ka + 16 * (2 * j + 1)


def main():
    rng = np.random.default_rng(0)
    # 1. the register FFT is a DFT
    x = rng.standard_normal(256) + 1j * rng.standard_normal(256)
    X = fwd256(x)
    ref = np.fft.fft(x)
    ka, kb = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
    assert np.allclose(X, ref[ka + 16 * kb]), "forward mapping"
    assert np.allclose(inv256(X) / 256, x), "inverse mapping"
    # slot layout is a bijection and freq_of_f inverts it
    seen = set()
    for a in range(16):
        for b in range(16):
            f, lane = slot_of(a, b)
            assert freq_of_f(f)[lane] == a + 16 * b
            seen.add((f, lane))
    assert len(seen) == 256

    # 2. whole pipeline on a small torus: H x W world, kernel 31x31 D4-symmetric, band-local pairing
    H, Wd, Lp = 48, 300, 8
    K = rng.random((31, 31))
    K = K + K[::-1] + K[:, ::-1] + K[::-1, ::-1]
    K = K + K.T
    K /= K.sum()
    A = rng.random((H, Wd))
    # reference: periodic cross-correlation
    U = np.zeros((H, Wd))
    for dy in range(-R, R + 1):
        for dx in range(-R, R + 1):
            U += K[R + dy, R + dx] * np.roll(np.roll(A, -dy, 0), -dx, 1)
    # framed state
    F = 16
    fr = np.pad(A, F, mode="wrap")
    # tap spectra per dy >= 0 (real): Khat[dy][k] = sum_dx K[R+dy][R+dx] cos(2 pi k dx / N) / N
    k = np.arange(N)
    Khat = np.array([[sum(K[R + dy, R + dx] * np.cos(2 * np.pi * kk * dx / N) for dx in range(-R, R + 1)) / N for kk in k]
                     for dy in range(R + 1)])
    out = np.full((H, Wd), np.nan)
    nstrips = -(-Wd // V)
    nbands = -(-H // (2 * Lp))
    for s in range(nstrips):
        x0 = s * V - R          # world column of sample n = 0
        for b in range(nbands):
            y0 = b * 2 * Lp
            # spectra rows zr = 0 .. Lp + 2R - 1: image rows (y0 + zr - R, y0 + Lp + zr - R)
            Z = []
            for zr in range(Lp + 2 * R):
                ra, rb = y0 + zr - R, y0 + Lp + zr - R
                rowa = np.array([fr[min(ra + F, H + 2 * F - 1), c + F] if c + F < Wd + 2 * F else 0.0 for c in range(x0, x0 + N)])
                rowb = np.array([fr[min(rb + F, H + 2 * F - 1), c + F] if c + F < Wd + 2 * F else 0.0 for c in range(x0, x0 + N)])
                Z.append(fwd256(rowa + 1j * rowb))
            Z = np.array(Z)     # [zr][ka][kb]
            for p in range(Lp):
                acc = np.zeros((16, 16), complex)
                for dy in range(-R, R + 1):
                    acc += Khat[abs(dy)][ka + 16 * kb] * Z[p + dy + R]
                y = inv256(acc)
                for n in range(R, N - R):
                    xx = x0 + n
                    if xx >= Wd:
                        continue
                    if y0 + p < H:
                        out[y0 + p, xx] = y[n].real
                    if y0 + Lp + p < H:
                        out[y0 + Lp + p, xx] = y[n].imag
    assert not np.isnan(out).any()
    err = np.abs(out - U).max()
    print(f"pipeline max err {err:.2e}")
    assert err < 1e-12
    print("march_proto OK")


if __name__ == "__main__":
    main()
