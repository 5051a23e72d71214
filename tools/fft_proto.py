"""numpy prototype of the in-place radix-4 DIF (forward) / DIT (inverse) FFT pair used by lenia_fft.cu,
and of the row-FFT x column-direct hybrid convolution with row pairing.  Validates the index math on CPU."""
import numpy as np


def digit_reverse4(N):
    L = int(round(np.log(N) / np.log(4)))
    assert 4 ** L == N
    idx = np.arange(N)
    rev = np.zeros(N, dtype=np.int64)
    for d in range(L):
        rev = rev * 4 + ((idx >> (2 * d)) & 3)
    return rev  # rev[n] = digit-reversed n


def fft_dif4_inplace(x, inverse=False):
    """In-place radix-4 decimation-in-frequency.  natural order in -> base-4 digit-reversed order out."""
    x = x.astype(np.complex128).copy()
    N = x.shape[-1]
    sgn = 1.0 if inverse else -1.0
    span = N // 4  # distance between butterfly legs
    while span >= 1:
        # butterflies: group base g (multiple of 4*span), offset k in [0, span)
        for j in range(N // 4):
            k = j % span
            g = (j // span) * 4 * span
            i0 = g + k
            a, b, c, d = x[..., i0], x[..., i0 + span], x[..., i0 + 2 * span], x[..., i0 + 3 * span]
            t0, t1, t2, t3 = a + c, a - c, b + d, (b - d) * (sgn * 1j)
            y0, y1, y2, y3 = t0 + t2, t1 + t3, t0 - t2, t1 - t3
            # DIF: twiddle AFTER the butterfly: output r gets w^(r*k), w = exp(sgn*2*pi*i/(4*span))
            w = np.exp(sgn * 2j * np.pi * k / (4 * span))
            x[..., i0] = y0
            x[..., i0 + span] = y1 * w
            x[..., i0 + 2 * span] = y2 * w**2
            x[..., i0 + 3 * span] = y3 * w**3
        span //= 4
    return x


def fft_dit4_inplace(x, inverse=True):
    """In-place radix-4 decimation-in-time.  digit-reversed order in -> natural order out."""
    x = x.astype(np.complex128).copy()
    N = x.shape[-1]
    sgn = 1.0 if inverse else -1.0
    span = 1
    while span < N:
        for j in range(N // 4):
            k = j % span
            g = (j // span) * 4 * span
            i0 = g + k
            w = np.exp(sgn * 2j * np.pi * k / (4 * span))
            a, b, c, d = x[..., i0], x[..., i0 + span] * w, x[..., i0 + 2 * span] * w**2, x[..., i0 + 3 * span] * w**3
            t0, t1, t2, t3 = a + c, a - c, b + d, (b - d) * (sgn * 1j)
            x[..., i0] = t0 + t2
            x[..., i0 + span] = t1 + t3
            x[..., i0 + 2 * span] = t0 - t2
            x[..., i0 + 3 * span] = t1 - t3
        span *= 4
    return x


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for N in (16, 64, 256, 1024):
        x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        X = np.fft.fft(x)
        rev = digit_reverse4(N)
        Y = fft_dif4_inplace(x)
        # Y[p] should be X[rev[p]]
        e1 = np.abs(Y - X[rev]).max()
        # which butterfly output order? check y1<->y3 convention
        back = fft_dit4_inplace(Y, inverse=True) / N
        e2 = np.abs(back - x).max()
        print(N, "dif vs fft (digit-reversed):", e1, " roundtrip:", e2)

    # hybrid convolution check: rows paired (r, r+D), x by FFT (periodic N == W), y direct with real K^ (symmetric taps)
    H, W, R = 32, 64, 5
    img = rng.random((H, W))
    ax = np.linspace(-1, 1, 2 * R + 1)
    gx, gy = np.meshgrid(ax, ax)
    K = np.exp(-((np.sqrt(gx**2 + gy**2) - 0.5) / 0.15) ** 2 / 2)
    K /= K.sum()
    ref = np.zeros_like(img)
    for dy in range(-R, R + 1):
        for dx in range(-R, R + 1):
            ref += K[dy + R, dx + R] * np.roll(img, (-dy, -dx), axis=(0, 1))
    # K^[dy][k] = sum_dx K[dy][dx] e^{-2 pi i k dx / N}: real because K[dy][dx] == K[dy][-dx]
    Khat = np.zeros((2 * R + 1, W))
    for dy in range(2 * R + 1):
        row = np.zeros(W)
        for dx in range(-R, R + 1):
            row[dx % W] += K[dy, dx + R]
        f = np.fft.fft(row)
        assert np.abs(f.imag).max() < 1e-12
        Khat[dy] = f.real
    rev = digit_reverse4(W)
    D = H // 2
    Z = fft_dif4_inplace(img[:D] + 1j * img[D:])          # (D, W) digit-reversed spectra of row pairs (r, r+D)
    Kp = Khat[:, rev] / W                                  # same permutation, 1/N folded in
    out = np.zeros((D, W), dtype=np.complex128)
    for y in range(D):
        for dy in range(-R, R + 1):
            # cross-correlation: out[y] += K[dy] * in[y+dy]; multiplying by the spectrum of the symmetric row
            r = (y + dy)
            if 0 <= r < D:
                zr = Z[r]
            elif r < 0:
                # row r (<0) pairs: image rows (r mod H, r+D mod H) = (r+H, r+D): that is Z[r+D] with re/im swapped roles:
                # Z_{r} = In[r] + i In[r+D];  for r<0: In[r+H] + i In[r+D] = imag-row of Z[r+D] as real, real-row as imag
                zr = None
            else:
                zr = None
            if zr is None:
                rr = r % H
                a = np.fft.fft(img[rr])[rev]
                b = np.fft.fft(img[(rr + D) % H])[rev]
                zr = a + 1j * b
            out[y] += Kp[dy + R] * zr
    res = fft_dit4_inplace(out, inverse=True)
    got = np.concatenate([res.real, res.imag], axis=0)
    print("hybrid conv max err:", np.abs(got - ref).max())
