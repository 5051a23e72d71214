// Lenia-family step, FFT path (row FFT x column-direct hybrid), for sm_100a.
//
// The reference convolves by full-plane complex 2-D FFTs (pyca/automata/models/lenia/lenia.py:453-469), a
// global transform that cannot be tiled or sharded.  This path keeps the FFT in ONE dimension only:
//   out[y][:] = sum_dy  K[dy][:] (*)_x  in[y+dy][:]
// and since every kernel row K[dy][.] is symmetric in dx (D4-symmetric kernels), its spectrum Khat[dy][k] is
// REAL.  So, per frequency k, the column direction is a 31-tap real-coefficient FIR over spectra of rows:
//   Out^[y][k] = sum_dy Khat[dy][k] * In^[y+dy][k].
// Real coefficients also let two image rows ride in one complex FFT: Z_r = FFT_x(in[r] + i*in[r+D]), D = H/2;
// the FIR over r and the inverse FFT keep real and imaginary parts separate, so out[y] + i*out[y+D] pops out
// with no real-FFT post-processing.  Cost ~31 FMA (FIR) + ~2x30 instructions (two radix-4 FFTs) per cell instead
// of 496 FMA for the folded direct path; the column direction stays local (15 ghost rows), so row slabs and
// the periodic frame work exactly as for the direct path.
//
//   kernel A  lenia_fft_rows_kernel   one CTA per row pair (r, r+D), r in [-15, D+15): in-place radix-4
//             decimation-in-frequency FFT in shared memory (natural order in, base-4 digit-reversed out),
//             spectrum written to a workspace in HBM.
//   kernel B  lenia_fft_cols_kernel   one CTA per TH row pairs x segment: FIR along rows straight from the
//             workspace (L2) into registers, spectra parked in shared memory, in-place radix-4
//             decimation-in-time inverse FFT (digit-reversed in, natural out; 1/N folded into Khat), then
//             growth + weighted source sum + Euler step/clamp (or affinity) with the periodic frame.
// The x direction uses the exact periodic transform when W == N (N in {256, 1024, 4096}), otherwise overlap-save
// segments of N = 1024 with N-30 valid outputs read from the framed layout.
#include "common.cuh"

namespace b200ca {

constexpr int F = B200CA_FRAME;
constexpr int R = 15;
constexpr int KROWS = R + 1;

struct FftGeom {
    int N, logN4;      // transform length (power of 4)
    int nseg, V;       // segments per row, valid outputs per segment
    int periodic;      // W == N: exact circular convolution, no halo in x
    int D;             // H / 2 (row pairing distance)
    int zrows;         // D + 2R spectra rows per plane
};

// complex helpers on float2
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
__device__ __forceinline__ float2 cmulc(float2 a, float2 w) { return make_float2(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y); }  // a*conj(w)

// One in-place radix-4 DIF butterfly (forward, e^{-i...}) on legs i0 + {0,1,2,3}*span; tw = exp(-2 pi i m / N) table.
__device__ __forceinline__ void dif4(float2 &a, float2 &b, float2 &c, float2 &d, float2 w1, float2 w2, float2 w3) {
    const float2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d);
    const float2 bd = csub(b, d);
    const float2 t3 = make_float2(bd.y, -bd.x);  // (b-d) * (-i)
    a = cadd(t0, t2);
    b = cmul(cadd(t1, t3), w1);
    c = cmul(csub(t0, t2), w2);
    d = cmul(csub(t1, t3), w3);
}
// One in-place radix-4 DIT butterfly (inverse, e^{+i...}): twiddles (conjugated table entries) applied first.
__device__ __forceinline__ void dit4_inv(float2 &a, float2 &b, float2 &c, float2 &d, float2 w1, float2 w2, float2 w3) {
    b = cmulc(b, w1);
    c = cmulc(c, w2);
    d = cmulc(d, w3);
    const float2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d);
    const float2 bd = csub(b, d);
    const float2 t3 = make_float2(-bd.y, bd.x);  // (b-d) * (+i)
    a = cadd(t0, t2);
    b = cadd(t1, t3);
    c = csub(t0, t2);
    d = csub(t1, t3);
}

// shared-memory index swizzle for float2 arrays: keeps the radix-4 stages with span 4 and 1 (and the linear
// passes) free of bank conflicts.  Only the low 4 bits change, so 16-element aligned blocks stay blocks.
__device__ __forceinline__ int swz(int i) { return i ^ (5 * ((i >> 4) & 3)); }

// ---- kernel A: forward row FFTs -------------------------------------------------------------------
// grid (zrows, nseg, B*C), block N/4 threads, smem 2N float2 (data + twiddles)
__global__ void __launch_bounds__(1024) lenia_fft_rows_kernel(const float *__restrict__ state, float2 *__restrict__ Z,
                                                               const float2 *__restrict__ twg, FftGeom g, int H, int W, int pitch) {
    extern __shared__ __align__(16) float2 fsm[];
    float2 *buf = fsm;           // [N]
    float2 *tw = fsm + g.N;      // [N]
    const int N = g.N, T = N >> 2, tid = threadIdx.x;
    const int zr = blockIdx.x, seg = blockIdx.y, plane = blockIdx.z;
    for (int t = tid; t < N; t += T) tw[t] = __ldg(twg + t);
    // image rows (framed): r = zr - R (may be negative: frame rows) and r + D
    const float *rowA = state + ((size_t)plane * (H + 2 * F) + (zr - R + F)) * pitch;
    const float *rowB = rowA + (size_t)g.D * pitch;
    const int xoff = g.periodic ? F : seg * g.V - R + F;  // framed column of sample n = 0
    // first stage straight from global memory (span = N/4)
    {
        float2 v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int col = xoff + tid + r * T;
            const bool ok = col < pitch;
            v[r] = make_float2(ok ? __ldg(rowA + col) : 0.f, ok ? __ldg(rowB + col) : 0.f);
        }
        __syncthreads();  // twiddles visible
        const int m = tid;  // k = tid, span = T: twiddle index k * N/(4*span) = k
        dif4(v[0], v[1], v[2], v[3], tw[m], tw[2 * m], tw[3 * m]);
#pragma unroll
        for (int r = 0; r < 4; ++r) buf[swz(tid + r * T)] = v[r];
    }
    // middle stages in shared memory
    for (int span = T >> 2; span > 1; span >>= 2) {
        __syncthreads();
        const int k = tid & (span - 1);
        const int i0 = ((tid - k) << 2) + k;
        const int m = k * (T / span);
        const int p0 = swz(i0), p1 = swz(i0 + span), p2 = swz(i0 + 2 * span), p3 = swz(i0 + 3 * span);
        float2 a = buf[p0], b = buf[p1], c = buf[p2], d = buf[p3];
        dif4(a, b, c, d, tw[m], tw[2 * m], tw[3 * m]);
        buf[p0] = a; buf[p1] = b; buf[p2] = c; buf[p3] = d;
    }
    // last stage (span = 1, twiddles = 1) straight to global memory: 4 consecutive outputs per thread
    __syncthreads();
    {
        const int i0 = tid << 2;
        float2 a = buf[swz(i0)], b = buf[swz(i0 + 1)], c = buf[swz(i0 + 2)], d = buf[swz(i0 + 3)];
        const float2 one = make_float2(1.f, 0.f);
        dif4(a, b, c, d, one, one, one);
        float4 *out = reinterpret_cast<float4 *>(Z + (((size_t)plane * g.zrows + zr) * g.nseg + seg) * N + i0);
        out[0] = make_float4(a.x, a.y, b.x, b.y);
        out[1] = make_float4(c.x, c.y, d.x, d.y);
    }
}

// ---- kernel B1: column FIR on the spectra ------------------------------------------------------------
// Out^[pair][y][k] = sum_dy Khat[pair][|dy|][k] * Z[src][y + dy][k].  One frequency per thread, FT output row pairs
// per thread; every spectrum value is loaded once per thread and fanned out to the <= 31 outputs it feeds.
struct FftFirArgs {
    const float2 *Z;
    float2 *O;
    const float *kfft;   // (B, NS, C, KROWS, N) real spectra, digit-reversed order, scaled by 1/N
    FftGeom g;
    int B, C, NS, k_mult;
};

constexpr int FIR_THREADS = 64;
template <int FT>
__global__ void __launch_bounds__(FIR_THREADS) lenia_fft_fir_kernel(const FftFirArgs a) {
    const int N = a.g.N;
    const int chunks = N / FIR_THREADS;
    const int k = (blockIdx.x % chunks) * FIR_THREADS + threadIdx.x;
    const int seg = blockIdx.x / chunks;
    const int y0 = blockIdx.y * FT;
    const int pair = blockIdx.z;  // (b * NS + i) * C + j
    const int j = pair % a.C, i = (pair / a.C) % a.NS, b = pair / (a.C * a.NS);
    (void)j;
    const int src_plane = b * a.C + i / a.k_mult;
    const float *kf = a.kfft + (size_t)pair * KROWS * N + k;
    float kt[KROWS];
#pragma unroll
    for (int dy = 0; dy < KROWS; ++dy) kt[dy] = __ldg(kf + dy * N);
    const size_t zstride = (size_t)a.g.nseg * N;
    const float2 *zp = a.Z + ((size_t)src_plane * a.g.zrows * a.g.nseg + seg) * N + k;
    float2 acc[FT];
#pragma unroll
    for (int y = 0; y < FT; ++y) acc[y] = make_float2(0.f, 0.f);
#pragma unroll
    for (int rr = 0; rr < FT + 2 * R; ++rr) {
        const int zr = min(y0 + rr, a.g.zrows - 1);  // spectra row zr holds image rows (zr - R, zr - R + D)
        const float2 z = __ldg(zp + (size_t)zr * zstride);
#pragma unroll
        for (int y = 0; y < FT; ++y) {
            const int dy = rr - R - y;
            if (dy >= -R && dy <= R) {
                const float w = kt[dy < 0 ? -dy : dy];
                acc[y].x = fmaf(w, z.x, acc[y].x);
                acc[y].y = fmaf(w, z.y, acc[y].y);
            }
        }
    }
    float2 *op = a.O + ((size_t)pair * a.g.D * a.g.nseg + seg) * N + k;
#pragma unroll
    for (int y = 0; y < FT; ++y)
        if (y0 + y < a.g.D) op[(size_t)(y0 + y) * zstride] = acc[y];
}

// ---- kernel B2: inverse row FFT + growth + source sum + Euler/clamp -------------------------------------
struct FftInvArgs {
    const float *state_in;
    float *out;
    const float2 *O;
    const float2 *twg;
    const float *mu, *sigma, *weights;
    FftGeom g;
    int B, C, NS, H, W, pitch;
    float dt;
    int mode, row_wrap;
};

__device__ __forceinline__ void fft_store_mirrors(float *__restrict__ plane, int H, int W, int pitch, int y, int x, float v, bool row_wrap) {
    const int fy = y + F, fx = x + F;
    int ys[3], xs[3], ny = 0, nx = 0;
    ys[ny++] = fy;
    if (row_wrap) {
        if (y < F) ys[ny++] = fy + H;
        if (y >= H - F) ys[ny++] = fy - H;
    }
    xs[nx++] = fx;
    if (x < F) xs[nx++] = fx + W;
    if (x >= W - F) xs[nx++] = fx - W;
    for (int a = 0; a < ny; ++a)
        for (int b = 0; b < nx; ++b) plane[(size_t)ys[a] * pitch + xs[b]] = v;
}

// grid (D, nseg, B*C), block N/4 threads
__global__ void __launch_bounds__(1024) lenia_fft_inv_kernel(const FftInvArgs a) {
    extern __shared__ __align__(16) float2 ism[];
    const int N = a.g.N, T = N >> 2, tid = threadIdx.x;
    float2 *buf = ism;
    float2 *tw = ism + N;
    const int yy = blockIdx.x, seg = blockIdx.y;
    const int b = blockIdx.z / a.C, j = blockIdx.z % a.C;
    for (int t = tid; t < N; t += T) tw[t] = __ldg(a.twg + t);
    float2 dx[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) dx[r] = make_float2(0.f, 0.f);
    for (int i = 0; i < a.NS; ++i) {
        const int pidx = (b * a.NS + i) * a.C + j;
        const float2 *op = a.O + (((size_t)pidx * a.g.D + yy) * a.g.nseg + seg) * N;
        // first DIT stage (span = 1, no twiddles) straight from global memory: 4 consecutive inputs per thread
        {
            const float4 *in4 = reinterpret_cast<const float4 *>(op + (tid << 2));
            const float4 lo = __ldg(in4), hi = __ldg(in4 + 1);
            float2 p = make_float2(lo.x, lo.y), q = make_float2(lo.z, lo.w), r = make_float2(hi.x, hi.y), s = make_float2(hi.z, hi.w);
            const float2 one = make_float2(1.f, 0.f);
            dit4_inv(p, q, r, s, one, one, one);
            __syncthreads();  // twiddles visible (i == 0); previous source done with buf
            const int i0 = tid << 2;
            buf[swz(i0)] = p; buf[swz(i0 + 1)] = q; buf[swz(i0 + 2)] = r; buf[swz(i0 + 3)] = s;
        }
        for (int span = 4; span < T; span <<= 2) {
            __syncthreads();
            const int k = tid & (span - 1);
            const int i0 = ((tid - k) << 2) + k;
            const int m = k * (T / span);
            const int p0 = swz(i0), p1 = swz(i0 + span), p2 = swz(i0 + 2 * span), p3 = swz(i0 + 3 * span);
            float2 p = buf[p0], q = buf[p1], r = buf[p2], s = buf[p3];
            dit4_inv(p, q, r, s, tw[m], tw[2 * m], tw[3 * m]);
            buf[p0] = p; buf[p1] = q; buf[p2] = r; buf[p3] = s;
        }
        __syncthreads();
        // last stage (span = N/4): results stay in registers: sample n = tid + r*T of rows yy (real) and yy+D (imag)
        float2 v[4];
        {
            v[0] = buf[swz(tid)]; v[1] = buf[swz(tid + T)]; v[2] = buf[swz(tid + 2 * T)]; v[3] = buf[swz(tid + 3 * T)];
            const int m = tid;
            dit4_inv(v[0], v[1], v[2], v[3], tw[m], tw[2 * m], tw[3 * m]);
        }
        // growth + weighted source sum (lenia.py:423-426, :444-446): 2*exp(-(u-mu)^2/(2 sigma^2)) - 1
        const float mu = __ldg(a.mu + pidx), sg = __ldg(a.sigma + pidx), wt = __ldg(a.weights + pidx);
        const float c2 = 1.4426950408889634f / (2.f * sg * sg);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float d0 = v[r].x - mu, d1 = v[r].y - mu;
            dx[r].x = fmaf(2.f * exp2f(-d0 * d0 * c2) - 1.f, wt, dx[r].x);
            dx[r].y = fmaf(2.f * exp2f(-d1 * d1 * c2) - 1.f, wt, dx[r].y);
        }
    }
    const size_t plane_off = (size_t)(b * a.C + j) * (a.H + 2 * F) * a.pitch;
    float *oplane = a.out + plane_off;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int n = tid + r * T;
        int x;
        if (a.g.periodic) {
            x = n;
        } else {
            if (n < R || n >= N - R) continue;
            x = seg * a.g.V + n - R;
            if (x >= a.W) continue;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int row = yy + h * a.g.D;
            const float dv = h ? dx[r].y : dx[r].x;
            float res = dv;
            if (a.mode == B200CA_MODE_LENIA) {
                const float sv = __ldg(a.state_in + plane_off + (size_t)(row + F) * a.pitch + x + F);
                res = fminf(fmaxf(sv + a.dt * dv, 0.f), 1.f);
            }
            const bool edge = (x < F) || (x >= a.W - F) || (a.row_wrap && (row < F || row >= a.H - F));
            if (edge)
                fft_store_mirrors(oplane, a.H, a.W, a.pitch, row, x, res, a.row_wrap != 0);
            else
                oplane[(size_t)(row + F) * a.pitch + x + F] = res;
        }
    }
}

// ---- plan helpers ---------------------------------------------------------------------------------
__global__ void lenia_fft_twiddle_kernel(float2 *tw, int N) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= N) return;
    double s, c;
    sincospi(-2.0 * (double)m / (double)N, &s, &c);
    tw[m] = make_float2((float)c, (float)s);
}

__device__ __forceinline__ int digit_reverse4(int p, int logN4) {
    int r = 0;
    for (int d = 0; d < logN4; ++d) {
        r = (r << 2) | (p & 3);
        p >>= 2;
    }
    return r;
}

// Khat[dy][p] = (1/N) * sum_dx K[R+dy][R+dx] cos(2 pi k dx / N),  k = digit_reverse4(p)   (double precision)
__global__ void lenia_fft_prepare_kernel(const float *__restrict__ K, float *__restrict__ kfft, int nker, int ks, int N, int logN4,
                                         int *__restrict__ asym_flag) {
    const int ker = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N || ker >= nker) return;
    const int k = digit_reverse4(p, logN4);
    const int off = (2 * R + 1 - ks) / 2;
    const float *kk = K + (size_t)ker * ks * ks;
    for (int dy = 0; dy < KROWS; ++dy) {
        double acc = 0.0;
        const int rp = R + dy - off, rm = R - dy - off;
        if (rp >= 0 && rp < ks && rm >= 0) {
            for (int c = 0; c < ks; ++c) {
                const int dx = c + off - R;
                const double up = kk[rp * ks + c], um = kk[rm * ks + c], mir = kk[rp * ks + (ks - 1 - c)];
                if (fabs(up - um) > 1e-6 * fmax(fabs(up), fabs(um)) + 1e-12 || fabs(up - mir) > 1e-6 * fmax(fabs(up), fabs(mir)) + 1e-12)
                    atomicExch(asym_flag, 1);
                acc += 0.5 * (up + um) * cospi(2.0 * (double)((long long)k * dx % N) / (double)N);
            }
        }
        kfft[((size_t)ker * KROWS + dy) * N + p] = (float)(acc / (double)N);
    }
}

static int fft_geom(FftGeom &g, int H, int W) {
    if (H % 2 != 0 || H < 2 * F) return B200CA_E_UNSUPPORTED;
    g.D = H / 2;
    g.zrows = g.D + 2 * R;
    if (W == 256 || W == 1024 || W == 4096) {
        g.N = W;
        g.periodic = 1;
        g.nseg = 1;
        g.V = W;
    } else {
        if (W < 64) return B200CA_E_UNSUPPORTED;
        g.N = W < 900 ? 256 : 1024;
        g.periodic = 0;
        g.V = g.N - 2 * R;
        g.nseg = ceil_div(W, g.V);
    }
    g.logN4 = g.N == 256 ? 4 : (g.N == 1024 ? 5 : 6);
    return 0;
}

// twiddle tables live for the lifetime of the library (one per device and N)
static float2 *g_tw[64][3] = {};
static int get_twiddles(int N, const float2 **out, cudaStream_t st) {
    int dev = 0;
    B200CA_CUDA(cudaGetDevice(&dev));
    const int slot = N == 256 ? 0 : (N == 1024 ? 1 : 2);
    if (!g_tw[dev & 63][slot]) {
        float2 *p = nullptr;
        B200CA_CUDA(cudaMalloc(&p, sizeof(float2) * N));
        lenia_fft_twiddle_kernel<<<ceil_div(N, 256), 256, 0, st>>>(p, N);
        B200CA_LAUNCH_CHECK("lenia_fft_twiddle_kernel");
        B200CA_CUDA(cudaStreamSynchronize(st));
        g_tw[dev & 63][slot] = p;
    }
    *out = g_tw[dev & 63][slot];
    return 0;
}

}  // namespace b200ca

using namespace b200ca;

extern "C" int b200ca_lenia_fft_length(int H, int W) {
    FftGeom g;
    return fft_geom(g, H, W) ? 0 : g.N;
}

extern "C" int64_t b200ca_lenia_fft_kernel_elems(int B, int C, int k_mult, int H, int W) {
    FftGeom g;
    if (fft_geom(g, H, W)) return -1;
    return (int64_t)B * C * k_mult * C * KROWS * g.N;
}

extern "C" int64_t b200ca_lenia_fft_workspace_elems(int B, int C, int k_mult, int H, int W) {
    FftGeom g;
    if (fft_geom(g, H, W)) return -1;
    return (int64_t)B * C * g.zrows * g.nseg * g.N * 2 + (int64_t)B * C * k_mult * C * g.D * g.nseg * g.N * 2;
}

extern "C" int b200ca_lenia_fft_prepare_kernel(const float *K, float *Kfft, int B, int C, int k_mult, int ks, int H, int W, void *stream) {
    B200CA_REQUIRE(K && Kfft && B >= 1 && C >= 1 && k_mult >= 1, "lenia_fft_prepare_kernel: bad arguments");
    FftGeom g;
    if (fft_geom(g, H, W) || ks < 1 || ks > 2 * R + 1 || (ks % 2) == 0) {
        set_error("lenia_fft_prepare_kernel: world %dx%d / k_size %d not supported by the FFT path", H, W, ks);
        return B200CA_E_UNSUPPORTED;
    }
    cudaStream_t st = as_stream(stream);
    int *flag = nullptr;
    B200CA_CUDA(cudaMalloc(&flag, sizeof(int)));
    int h_flag = 0;
    int rc = check_cuda(cudaMemsetAsync(flag, 0, sizeof(int), st), "memset");
    if (!rc) {
        const int nker = B * C * k_mult * C;
        dim3 grid(ceil_div(g.N, 128), nker);
        lenia_fft_prepare_kernel<<<grid, 128, 0, st>>>(K, Kfft, nker, ks, g.N, g.logN4, flag);
        rc = check_cuda(cudaPeekAtLastError(), "lenia_fft_prepare_kernel");
    }
    if (!rc) rc = check_cuda(cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st), "D2H flag");
    if (!rc) rc = check_cuda(cudaStreamSynchronize(st), "sync");
    cudaFree(flag);
    if (rc) return rc;
    if (h_flag) {
        set_error("lenia_fft_prepare_kernel: kernel rows are not symmetric (the FFT path needs K[dy][dx]==K[-dy][dx]==K[dy][-dx])");
        return B200CA_E_UNSUPPORTED;
    }
    return 0;
}

extern "C" int b200ca_lenia_step_fft(const float *state_in, float *state_out, const float *Kfft, const float *mu, const float *sigma,
                                     const float *weights, float *workspace, int B, int C, int k_mult, int H, int W, float dt, int mode,
                                     int row_wrap, void *stream) {
    B200CA_REQUIRE(state_in && state_out && Kfft && mu && sigma && weights && workspace, "lenia_step_fft: null pointer");
    B200CA_REQUIRE(state_in != state_out, "lenia_step_fft: in-place update is not supported");
    B200CA_REQUIRE(mode == B200CA_MODE_LENIA || mode == B200CA_MODE_AFF, "lenia_step_fft: bad mode %d", mode);
    B200CA_REQUIRE((long long)B * C * C * k_mult <= 65535, "lenia_step_fft: too many kernel pairs");
    FftGeom g;
    if (fft_geom(g, H, W)) {
        set_error("lenia_step_fft: world %dx%d not supported by the FFT path (H even >= 32, W >= 64)", H, W);
        return B200CA_E_UNSUPPORTED;
    }
    cudaStream_t st = as_stream(stream);
    const float2 *tw = nullptr;
    int rc = get_twiddles(g.N, &tw, st);
    if (rc) return rc;
    const int pitch = b200ca_frame_pitch(W);
    const int NS = C * k_mult;
    float2 *Z = reinterpret_cast<float2 *>(workspace);
    float2 *O = Z + (size_t)B * C * g.zrows * g.nseg * g.N;
    const size_t smem = (size_t)2 * g.N * sizeof(float2);
    static bool attr_done[64] = {false};
    int dev = 0;
    B200CA_CUDA(cudaGetDevice(&dev));
    if (!attr_done[dev & 63]) {
        B200CA_CUDA(cudaFuncSetAttribute(lenia_fft_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        B200CA_CUDA(cudaFuncSetAttribute(lenia_fft_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_done[dev & 63] = true;
    }
    {
        dim3 grid(g.zrows, g.nseg, B * C);
        lenia_fft_rows_kernel<<<grid, g.N / 4, smem, st>>>(state_in, Z, tw, g, H, W, pitch);
        B200CA_LAUNCH_CHECK("lenia_fft_rows_kernel");
    }
    {
        FftFirArgs a;
        a.Z = Z;
        a.O = O;
        a.kfft = Kfft;
        a.g = g;
        a.B = B;
        a.C = C;
        a.NS = NS;
        a.k_mult = k_mult;
        constexpr int FT = 32;
        dim3 grid((g.N / FIR_THREADS) * g.nseg, ceil_div(g.D, FT), B * NS * C);
        lenia_fft_fir_kernel<FT><<<grid, FIR_THREADS, 0, st>>>(a);
        B200CA_LAUNCH_CHECK("lenia_fft_fir_kernel");
    }
    {
        FftInvArgs a;
        a.state_in = state_in;
        a.out = state_out;
        a.O = O;
        a.twg = tw;
        a.mu = mu;
        a.sigma = sigma;
        a.weights = weights;
        a.g = g;
        a.B = B;
        a.C = C;
        a.NS = NS;
        a.H = H;
        a.W = W;
        a.pitch = pitch;
        a.dt = dt;
        a.mode = mode;
        a.row_wrap = row_wrap;
        dim3 grid(g.D, g.nseg, B * C);
        lenia_fft_inv_kernel<<<grid, g.N / 4, smem, st>>>(a);
        B200CA_LAUNCH_CHECK("lenia_fft_inv_kernel");
    }
    return 0;
}
