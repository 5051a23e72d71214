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
//   kernel A  lenia_fft_rows_kernel  one CTA per (two) row pairs (r, r+D), r in [-15, D+15): in-place radix-4
//             decimation-in-frequency FFT in shared memory (natural order in, base-4 digit-reversed out; first
//             and last stage fused with the global loads/stores), spectra Z written to a workspace.
//   kernel B1 lenia_fft_fir_kernel   the column FIR: one frequency per thread, 32 output rows per thread, every
//             spectrum value loaded once and fanned out with packed FFMA2; writes the filtered spectra O.
//   kernel B2 lenia_fft_inv_kernel   one CTA per row pair: in-place radix-4 decimation-in-time inverse FFT
//             (digit-reversed in, natural out; 1/N folded into Khat), then growth + weighted source sum +
//             Euler step/clamp (or affinity) and the periodic frame of the output.
// The x direction uses the exact periodic transform when W == N (N in {256, 1024}), otherwise overlap-save
// segments (N = 1024, or 256 for narrow worlds) with N-30 valid outputs read from the framed layout.
#include <stdlib.h>

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
    // sub-range of spectra rows one row-FFT launch covers: zr = zr_begin + i (+ zr_gap once i >= zr_gap_at), i < zr_count
    int zr_begin, zr_count, zr_gap_at, zr_gap;
};

// complex helpers on float2
// (packed fp32x2 instructions of sm_100: one FADD2 / FFMA2 per complex add / real-times-complex FMA)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
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

// Programmatic dependent launch: the three kernels of a step (and consecutive steps) are chained on one stream.
// Each kernel does its input-independent prologue, then waits for the previous grid (griddepcontrol.wait) and
// immediately lets the next grid start ITS prologue (griddepcontrol.launch_dependents), which takes the launch
// latency off the critical path of small worlds.
// shared-memory index swizzle for float2 arrays: keeps the radix-4 stages with span 4 and 1 (and the linear
// passes) free of bank conflicts.  Only the low 4 bits change, so 16-element aligned blocks stay blocks.
__device__ __forceinline__ int swz(int i) { return i ^ (5 * ((i >> 4) & 3)); }

// Twiddles: one contiguous table per radix-4 stage so that consecutive butterflies read consecutive entries:
//   tws[(span - 1) + r*span + k] = exp(-2 pi i (r+1) k / (4 span)),  r = 0..2, k < span   (N - 1 entries in total)
// read with __ldg (8 KB, L1-resident) -- no shared-memory copy, no bank conflicts.
__device__ __forceinline__ void load_tw(const float2 *__restrict__ tws, int span, int k, float2 &w1, float2 &w2, float2 &w3) {
    const float2 *t = tws + (span - 1) + k;
    w1 = __ldg(t);
    w2 = __ldg(t + span);
    w3 = __ldg(t + 2 * span);
}


// ---- kernel A: forward row FFTs -------------------------------------------------------------------
// grid (ceil(zrows / FFT_RPC), nseg, B*C), block N/4 threads, smem FFT_RPC * N float2
// SPLIT: adjacent frequency pairs are stored as (re0, re1, im0, im1) -- the operand layout of the fused FIR kernel
template <int FFT_RPC, int N, bool SPLIT>
__global__ void __launch_bounds__(N / 4) lenia_fft_rows_kernel(const float *__restrict__ state, float2 *__restrict__ Z,
                                                              const float2 *__restrict__ tws, FftGeom g, int H, int W, int pitch) {
    extern __shared__ __align__(16) float2 fsm[];
    constexpr int T = N >> 2;
    const int tid = threadIdx.x;
    const int seg = blockIdx.y, plane = blockIdx.z;
    const int xoff = g.periodic ? F : seg * g.V - R + F;  // framed column of sample n = 0
    int zr[FFT_RPC];
#pragma unroll
    for (int q = 0; q < FFT_RPC; ++q) {
        const int i = min((int)blockIdx.x * FFT_RPC + q, g.zr_count - 1);
        zr[q] = g.zr_begin + i + (i >= g.zr_gap_at ? g.zr_gap : 0);
    }
    // twiddles of every stage are fetched up front so their latency overlaps the first global loads instead of sitting
    // on the critical path of each stage (stage s: span = T >> 2s; the last stage, span 1, needs none)
    float2 tw1[5], tw2[5], tw3[5];
    {
        int sp = T;
#pragma unroll
        for (int st = 0; st < 5; ++st) {
            if (sp > 1) load_tw(tws, sp, tid & (sp - 1), tw1[st], tw2[st], tw3[st]);
            sp >>= 2;
        }
    }
    pdl_wait_then_release();  // the state is the previous step's output
    // first stage straight from global memory (span = N/4, k = tid)
    {
        const float2 w1 = tw1[0], w2 = tw2[0], w3 = tw3[0];
#pragma unroll
        for (int q = 0; q < FFT_RPC; ++q) {
            // image rows (framed): r = zr - R (may be negative: frame rows) and r + D
            const float *rowA = state + ((size_t)plane * (H + 2 * F) + (zr[q] - R + F)) * pitch;
            const float *rowB = rowA + (size_t)g.D * pitch;
            float2 v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int col = xoff + tid + r * T;
                const bool ok = col < pitch;
                v[r] = make_float2(ok ? __ldg(rowA + col) : 0.f, ok ? __ldg(rowB + col) : 0.f);
            }
            dif4(v[0], v[1], v[2], v[3], w1, w2, w3);
            float2 *buf = fsm + q * N;
#pragma unroll
            for (int r = 0; r < 4; ++r) buf[swz(tid + r * T)] = v[r];
        }
    }
    // middle stages in shared memory
    int stg = 1;
#pragma unroll
    for (int span = T >> 2; span > 1; span >>= 2, ++stg) {
        __syncthreads();
        const int k = tid & (span - 1);
        const int i0 = ((tid - k) << 2) + k;
        const float2 w1 = tw1[stg], w2 = tw2[stg], w3 = tw3[stg];
        // the swizzle only looks at index bits 4-5, which adding a multiple of 64 does not touch
        const int p0 = swz(i0);
        const int p1 = span >= 64 ? p0 + span : swz(i0 + span);
        const int p2 = span >= 64 ? p0 + 2 * span : swz(i0 + 2 * span);
        const int p3 = span >= 64 ? p0 + 3 * span : swz(i0 + 3 * span);
#pragma unroll
        for (int q = 0; q < FFT_RPC; ++q) {
            float2 *buf = fsm + q * N;
            float2 a = buf[p0], b = buf[p1], c = buf[p2], d = buf[p3];
            dif4(a, b, c, d, w1, w2, w3);
            buf[p0] = a; buf[p1] = b; buf[p2] = c; buf[p3] = d;
        }
    }
    // last stage (span = 1, twiddles = 1) straight to global memory: 4 consecutive outputs per thread
    __syncthreads();
#pragma unroll
    for (int q = 0; q < FFT_RPC; ++q) {
        if ((int)blockIdx.x * FFT_RPC + q >= g.zr_count) break;
        const float2 *buf = fsm + q * N;
        const int i0 = tid << 2;
        float2 a = buf[swz(i0)], b = buf[swz(i0 + 1)], c = buf[swz(i0 + 2)], d = buf[swz(i0 + 3)];
        const float2 one = make_float2(1.f, 0.f);
        dif4(a, b, c, d, one, one, one);
        float4 *out = reinterpret_cast<float4 *>(Z + (((size_t)plane * g.zrows + zr[q]) * g.nseg + seg) * N + i0);
        out[0] = SPLIT ? make_float4(a.x, b.x, a.y, b.y) : make_float4(a.x, a.y, b.x, b.y);
        out[1] = SPLIT ? make_float4(c.x, d.x, c.y, d.y) : make_float4(c.x, c.y, d.x, d.y);
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
    float2 kt[KROWS];  // (w, w): real tap duplicated for the packed FMA
#pragma unroll
    for (int dy = 0; dy < KROWS; ++dy) {
        const float w = __ldg(kf + dy * N);
        kt[dy] = make_float2(w, w);
    }
    pdl_wait_then_release();  // taps are static; the spectra come from the row-FFT kernel
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
            if (dy >= -R && dy <= R) acc[y] = __ffma2_rn(kt[dy < 0 ? -dy : dy], z, acc[y]);
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
    const float2 *Z;     // fused FIR + inverse kernel only: the row spectra and the real tap spectra
    const float *kfft;
    int k_mult;
    const float2 *tws;
    const float *mu, *sigma, *weights;
    FftGeom g;
    int B, C, NS, H, W, pitch;
    float dt;
    int mode, row_wrap;
};

// writes the periodic images (not the cell itself) of interior cell (y, x) into the frame; rare, kept out of line
__device__ __noinline__ void fft_store_mirrors(float *__restrict__ plane, int H, int W, int pitch, int y, int x, float v, bool row_wrap) {
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
        for (int b = 0; b < nx; ++b)
            if (a | b) plane[(size_t)ys[a] * pitch + xs[b]] = v;
}

// grid (ceil(D / FFT_RPC), nseg, B*C), block N/4 threads, smem FFT_RPC * N float2
template <int FFT_RPC, int N>
__global__ void __launch_bounds__(N / 4, (FFT_RPC == 1 && N == 1024) ? 4 : 1) lenia_fft_inv_kernel(const FftInvArgs a) {
    extern __shared__ __align__(16) float2 ism[];
    constexpr int T = N >> 2;
    const int tid = threadIdx.x;
    const int seg = blockIdx.y;
    const int b = blockIdx.z / a.C, j = blockIdx.z % a.C;
    int yy[FFT_RPC];
#pragma unroll
    for (int q = 0; q < FFT_RPC; ++q) yy[q] = min((int)blockIdx.x * FFT_RPC + q, a.g.D - 1);
    float2 dx[FFT_RPC][4], sv[FFT_RPC][4];
#pragma unroll
    for (int q = 0; q < FFT_RPC; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r) dx[q][r] = sv[q][r] = make_float2(0.f, 0.f);
    const size_t plane_off = (size_t)(b * a.C + j) * (a.H + 2 * F) * a.pitch;
    pdl_wait_then_release();  // filtered spectra come from the FIR kernel
    for (int i = 0; i < a.NS; ++i) {
        const int pidx = (b * a.NS + i) * a.C + j;
        if (i > 0) __syncthreads();  // previous source done with the buffers
        // first DIT stage (span = 1, no twiddles) straight from global memory: 4 consecutive inputs per thread
#pragma unroll
        for (int q = 0; q < FFT_RPC; ++q) {
            const float2 *op = a.O + (((size_t)pidx * a.g.D + yy[q]) * a.g.nseg + seg) * N;
            const float4 *in4 = reinterpret_cast<const float4 *>(op + (tid << 2));
            const float4 lo = __ldg(in4), hi = __ldg(in4 + 1);
            float2 p = make_float2(lo.x, lo.y), qq = make_float2(lo.z, lo.w), r = make_float2(hi.x, hi.y), s = make_float2(hi.z, hi.w);
            const float2 one = make_float2(1.f, 0.f);
            dit4_inv(p, qq, r, s, one, one, one);
            float2 *buf = ism + q * N;
            const int i0 = tid << 2;
            buf[swz(i0)] = p; buf[swz(i0 + 1)] = qq; buf[swz(i0 + 2)] = r; buf[swz(i0 + 3)] = s;
        }
#pragma unroll
        for (int span = 4; span < T; span <<= 2) {
            __syncthreads();
            const int k = tid & (span - 1);
            const int i0 = ((tid - k) << 2) + k;
            float2 w1, w2, w3;  // (prefetching all stages' twiddles costs more in occupancy than it hides: measured)
            load_tw(a.tws, span, k, w1, w2, w3);
            const int p0 = swz(i0);
            const int p1 = span >= 64 ? p0 + span : swz(i0 + span);
            const int p2 = span >= 64 ? p0 + 2 * span : swz(i0 + 2 * span);
            const int p3 = span >= 64 ? p0 + 3 * span : swz(i0 + 3 * span);
#pragma unroll
            for (int q = 0; q < FFT_RPC; ++q) {
                float2 *buf = ism + q * N;
                float2 p = buf[p0], qq = buf[p1], r = buf[p2], s = buf[p3];
                dit4_inv(p, qq, r, s, w1, w2, w3);
                buf[p0] = p; buf[p1] = qq; buf[p2] = r; buf[p3] = s;
            }
        }
        __syncthreads();
        if (i == a.NS - 1 && a.mode == B200CA_MODE_LENIA) {
            // prefetch the state values of the Euler step; their latency hides behind the last stage + growth
#pragma unroll
            for (int q = 0; q < FFT_RPC; ++q)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int n = tid + r * T;
                    const int x = a.g.periodic ? n : min(max(seg * a.g.V + n - R, 0), a.W - 1);
                    const float *sp = a.state_in + plane_off + (size_t)(yy[q] + F) * a.pitch + x + F;
                    sv[q][r] = make_float2(__ldg(sp), __ldg(sp + (size_t)a.g.D * a.pitch));
                }
        }
        // last stage (span = N/4): results stay in registers: sample n = tid + r*T of rows yy (real) and yy+D (imag)
        const float mu = __ldg(a.mu + pidx), sg = __ldg(a.sigma + pidx), wt = __ldg(a.weights + pidx);
        const float c2 = 1.4426950408889634f / (2.f * sg * sg);
        float2 w1, w2, w3;
        load_tw(a.tws, T, tid, w1, w2, w3);
#pragma unroll
        for (int q = 0; q < FFT_RPC; ++q) {
            const float2 *buf = ism + q * N;
            float2 v[4];
            v[0] = buf[swz(tid)]; v[1] = buf[swz(tid + T)]; v[2] = buf[swz(tid + 2 * T)]; v[3] = buf[swz(tid + 3 * T)];
            dit4_inv(v[0], v[1], v[2], v[3], w1, w2, w3);
            // growth + weighted source sum (lenia.py:423-426, :444-446): 2*exp(-(u-mu)^2/(2 sigma^2)) - 1
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float d0 = v[r].x - mu, d1 = v[r].y - mu;
                dx[q][r].x = fmaf(2.f * exp2f(-d0 * d0 * c2) - 1.f, wt, dx[q][r].x);
                dx[q][r].y = fmaf(2.f * exp2f(-d1 * d1 * c2) - 1.f, wt, dx[q][r].y);
            }
        }
    }
    float *oplane = a.out + plane_off;
    const bool lenia = a.mode == B200CA_MODE_LENIA;
#pragma unroll
    for (int q = 0; q < FFT_RPC; ++q) {
        if ((int)blockIdx.x * FFT_RPC + q >= a.g.D) break;
        const int row0 = yy[q], row1 = yy[q] + a.g.D;
        float *rp0 = oplane + (size_t)(row0 + F) * a.pitch + F;
        float *rp1 = rp0 + (size_t)a.g.D * a.pitch;
        const bool re0 = a.row_wrap && (row0 < F || row0 >= a.H - F);
        const bool re1 = a.row_wrap && (row1 < F || row1 >= a.H - F);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int n = tid + r * T;
            int x = n;
            if (!a.g.periodic) {
                x = seg * a.g.V + n - R;
                if (n < R || n >= N - R || x >= a.W) continue;
            }
            float v0 = dx[q][r].x, v1 = dx[q][r].y;
            if (lenia) {
                v0 = fminf(fmaxf(sv[q][r].x + a.dt * v0, 0.f), 1.f);
                v1 = fminf(fmaxf(sv[q][r].y + a.dt * v1, 0.f), 1.f);
            }
            rp0[x] = v0;
            rp1[x] = v1;
            const bool xe = (x < F) || (x >= a.W - F);
            if (xe || re0) fft_store_mirrors(oplane, a.H, a.W, a.pitch, row0, x, v0, a.row_wrap != 0);
            if (xe || re1) fft_store_mirrors(oplane, a.H, a.W, a.pitch, row1, x, v1, a.row_wrap != 0);
        }
    }
}

// ---- radix-16 variants for N = 1024 (the production kernels; the radix-4 ones above serve N = 256) -------------
// Same five radix-4 DIF/DIT stages (same butterflies, same digit-reversed spectrum order), but executed two stages at a
// time on 16 values held in registers by one thread, so a row takes 64 threads and TWO shared-memory exchanges instead
// of four (tools/fft16_proto.py is the numpy model of this mapping and of the bank-conflict-free layouts):
//   forward  R16(span 256, 64) on idx t + 64 q  ->  exchange 1  ->  R16(16, 4) on idx 64 B + k3 + 4 q  ->  exchange 2
//            ->  R4(1) on idx 16 t + e  ->  128 contiguous bytes per thread to global memory
//   inverse  the mirror image; the last R16 leaves samples n = t + 64 q in registers for the growth epilogue.
constexpr int XB = 1152;  // float2 per row buffer: 1024 + up to 128 pad (idx + idx/16, or idx + 2*(idx/16) for split spectra)
// barrier among the 64 threads (two warps) that own one row of a multi-row CTA
__device__ __forceinline__ void row_sync(int rw) { asm volatile("bar.sync %0, 64;" ::"r"(rw + 1) : "memory"); }
__device__ __forceinline__ int swz_a(int i) { return i ^ (4 * ((i >> 6) & 3)); }

__device__ __forceinline__ void dit4_nt(float2 &a, float2 &b, float2 &c, float2 &d) {
    const float2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d);
    const float2 bd = csub(b, d);
    const float2 t3 = make_float2(-bd.y, bd.x);
    a = cadd(t0, t2);
    b = cadd(t1, t3);
    c = csub(t0, t2);
    d = csub(t1, t3);
}

// two fused DIT stages: span SB on legs 4 blk + r first, then span SA on legs (m, m+4, m+8, m+12) with twiddle index ka0 + kas*m.
// `tws` is the shared-memory copy of the twiddle tables (plain loads).
__device__ __forceinline__ void load_tw_s(const float2 *tws, int span, int k, float2 &w1, float2 &w2, float2 &w3) {
    const float2 *t = tws + (span - 1) + k;
    w1 = t[0];
    w2 = t[span];
    w3 = t[2 * span];
}
template <int SB, int SA>
__device__ __forceinline__ void r16_dit(float2 (&v)[16], const float2 *tws, int kb, int ka0, int kas) {
    float2 w1, w2, w3;
    load_tw_s(tws, SB, kb, w1, w2, w3);
#pragma unroll
    for (int blk = 0; blk < 4; ++blk) dit4_inv(v[4 * blk], v[4 * blk + 1], v[4 * blk + 2], v[4 * blk + 3], w1, w2, w3);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        load_tw_s(tws, SA, ka0 + kas * m, w1, w2, w3);
        dit4_inv(v[m], v[m + 4], v[m + 8], v[m + 12], w1, w2, w3);
    }
}

__device__ __forceinline__ void cp_async4(unsigned dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Stages the state values (rows yy and yy + D) under one thread's 16 samples n = t + 64 q into the row's (free)
// shared-memory buffer as float2 slot n, by cp.async: no registers held while the copies are in flight.
// Columns are clamped into the framed row, so samples that are never stored (segment halo, beyond W) read legal memory.
__device__ __forceinline__ void inv16_stage_state(const FftInvArgs &a, float2 *buf, size_t plane_off, int yy, int seg, int t) {
    const int xb = a.g.periodic ? t : seg * a.g.V - R + t;
    const float *sp0 = a.state_in + plane_off + (size_t)(yy + F) * a.pitch + F;
    const float *sp1 = sp0 + (size_t)a.g.D * a.pitch;
    const int xmax = a.W + F - 1;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(buf + t);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int x = min(xb + 64 * q, xmax);
        cp_async4(dst + 512 * q, sp0 + x);
        cp_async4(dst + 512 * q + 4, sp1 + x);
    }
}

// Euler step / clamp (or affinity) and the stores of one thread's 16 samples n = t + 64 q of rows yy and yy + D;
// `sv` is the row buffer holding the staged state (inv16_stage_state).
// EDGE = false: an inner segment of a row with no frame images -- no bounds or mirror logic at all.
template <bool EDGE>
__device__ __forceinline__ void inv16_store(const FftInvArgs &a, const float2 (&dx)[16], const float2 *sv, size_t plane_off, int yy, int seg,
                                            int t) {
    constexpr int N = 1024;
    float *oplane = a.out + plane_off;
    const bool lenia = a.mode == B200CA_MODE_LENIA;
    const int row0 = yy, row1 = yy + a.g.D;
    const int xb = a.g.periodic ? t : seg * a.g.V - R + t;  // x of sample q = 0
    float *rp0 = oplane + (size_t)(row0 + F) * a.pitch + F + xb;
    float *rp1 = rp0 + (size_t)a.g.D * a.pitch;
    const bool re0 = EDGE && a.row_wrap && (row0 < F || row0 >= a.H - F);
    const bool re1 = EDGE && a.row_wrap && (row1 < F || row1 >= a.H - F);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int n = t + 64 * q;
        const int x = xb + 64 * q;
        if (!a.g.periodic) {
            if ((q == 0 && n < R) || (q == 15 && n >= N - R)) continue;
            if (EDGE && x >= a.W) continue;
        }
        float v0 = dx[q].x, v1 = dx[q].y;
        if (lenia) {
            const float2 s = sv[n];
            v0 = fminf(fmaxf(s.x + a.dt * v0, 0.f), 1.f);
            v1 = fminf(fmaxf(s.y + a.dt * v1, 0.f), 1.f);
        }
        rp0[64 * q] = v0;
        rp1[64 * q] = v1;
        if (EDGE) {
            if (!re0 && !re1) {
                // the common edge case: only the x images of the two cells (right frame for x < F, left frame for x >= W - F)
                if (x < F) {
                    rp0[64 * q + a.W] = v0;
                    rp1[64 * q + a.W] = v1;
                } else if (x >= a.W - F) {
                    rp0[64 * q - a.W] = v0;
                    rp1[64 * q - a.W] = v1;
                }
            } else {
                const bool xe = (x < F) || (x >= a.W - F);
                if (xe || re0) fft_store_mirrors(oplane, a.H, a.W, a.pitch, row0, x, v0, a.row_wrap != 0);
                if (xe || re1) fft_store_mirrors(oplane, a.H, a.W, a.pitch, row1, x, v1, a.row_wrap != 0);
            }
        }
    }
}

// ---- fused kernel B (N = 1024): column FIR on the spectra + inverse row FFT + growth/Euler epilogue -----------------
// One CTA makes FT output row pairs of one segment of one destination plane.  Phase 1 is the FIR: two adjacent
// frequencies per thread, so the real taps (w_p0, w_p1) and the split spectra (re_p0, re_p1), (im_p0, im_p1) are
// natural fp32x2 operands; a spectrum row is read once per CTA -- (FT+30)/FT = 4.75x read amplification at FT = 8,
// served by L2 because neighbouring row blocks run at the same time.  The filtered spectra stay in shared memory,
// where phase 2 runs the radix-16 inverse transforms (4 rows per pass) and the epilogue; they never go to HBM.
// grid (ceil(D / FT), nseg, B*C), block 256 threads, smem FT * XB float2
#ifdef B200CA_FFT_TRACE
__device__ long long g_fft_trace[8 * 4096];
#define FFT_TRACE(slot)                                                                                         \
    do {                                                                                                        \
        const int cta_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);                        \
        if (threadIdx.x == 0 && cta_ < 4096) g_fft_trace[cta_ * 8 + (slot)] = clock64();                        \
    } while (0)
#else
#define FFT_TRACE(slot)
#endif

// NT = 256 threads: two frequency chunks per thread in the FIR, 4 rows per inverse pass, 2 CTAs per SM (large worlds).
// NT = 512 (worlds whose CTAs all fit the machine at one per SM, e.g. 1024^2: the step is then one wave and a CTA's serial
//      chain IS the kernel time): one chunk per thread, all FT = 8 rows inverted in one pass, and the state values of the
//      Euler step are staged into their own shared-memory region in the prologue -- before the dependency wait, they are
//      the previous step's output -- instead of behind the last FFT stages.
template <int FT, bool MULTI, int MINB, int NT = 256>
__global__ void __launch_bounds__(NT, MINB) lenia_fft_firinv_kernel(const FftInvArgs a) {
    extern __shared__ __align__(16) float2 fism[];
    constexpr int N = 1024, RG = NT / 64, NP = FT / RG, NCH = N / (2 * NT);
    constexpr bool EARLY = NT == 512;
    static_assert(FT % RG == 0 && NP >= 1, "FT must be a multiple of the row groups per pass");
    const int tid = threadIdx.x, t = tid & 63, rw = tid >> 6;
    const int seg = blockIdx.y;
    const int b = blockIdx.z / a.C, j = blockIdx.z % a.C;
    const int y0 = (int)blockIdx.x * FT;
    const int B4 = t >> 2, k3 = t & 3;
    const bool lenia = a.mode == B200CA_MODE_LENIA;
    float2 dx[MULTI ? NP : 1][16];
    if (MULTI) {
#pragma unroll
        for (int ps = 0; ps < NP; ++ps)
#pragma unroll
            for (int q = 0; q < 16; ++q) dx[ps][q] = make_float2(0.f, 0.f);
    }
    const size_t plane_off = (size_t)(b * a.C + j) * (a.H + 2 * F) * a.pitch;
    const size_t zstride = (size_t)a.g.nseg * N;
    // everything that does not depend on the row-FFT kernel is fetched before the dependency wait, so that under
    // programmatic dependent launch it overlaps the previous kernel: the twiddle tables (copied to shared memory:
    // the inverse stages then never wait on L2) and the taps of the first chunk
    FFT_TRACE(0);
    float2 *tw_s = fism + FT * XB;
    float2 *st_s = tw_s + N;  // EARLY: FT rows of 1024 staged state pairs
    if (EARLY && lenia) {
#pragma unroll
        for (int ps = 0; ps < NP; ++ps)
            inv16_stage_state(a, st_s + (ps * RG + rw) * N, plane_off, min(y0 + ps * RG + rw, a.g.D - 1), seg, t);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // twiddle tables -> shared memory by cp.async (16 bytes = two entries per copy): nothing waits for them before the
    // inverse transforms, so their latency hides behind the FIR instead of sitting in the prologue
    for (int e = tid; e < N / 2; e += NT)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(tw_s + 2 * e)), "l"(a.tws + 2 * e)
                     : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    float2 kt[KROWS];
    {
        const float *kf = a.kfft + (size_t)((b * a.NS) * a.C + j) * KROWS * N + 2 * tid;
#pragma unroll
        for (int dy = 0; dy < KROWS; ++dy) kt[dy] = __ldg(reinterpret_cast<const float2 *>(kf + dy * N));
    }
    FFT_TRACE(1);
    pdl_wait_then_release();  // the spectra come from the row-FFT kernel
    FFT_TRACE(2);
    for (int i = 0; i < a.NS; ++i) {
        const int pidx = (b * a.NS + i) * a.C + j;
        const int src_plane = b * a.C + i / a.k_mult;
        if (i > 0) __syncthreads();  // previous source done with the buffers
        // phase 1: FIR.  Out^[y][p] = sum_dy Khat[|dy|][p] * Z[y + dy][p] for p = p0, p0 + 1
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
            const int p0 = 2 * tid + 2 * NT * c;
            if (i > 0 || c > 0) {
                const float *kf = a.kfft + (size_t)pidx * KROWS * N + p0;
#pragma unroll
                for (int dy = 0; dy < KROWS; ++dy) kt[dy] = __ldg(reinterpret_cast<const float2 *>(kf + dy * N));
            }
            const float2 *zp = a.Z + (((size_t)src_plane * a.g.zrows + y0) * a.g.nseg + seg) * N + p0;
            float2 are[FT], aim[FT];
#pragma unroll
            for (int y = 0; y < FT; ++y) are[y] = aim[y] = make_float2(0.f, 0.f);
#pragma unroll
            for (int rr = 0; rr < FT + 2 * R; ++rr) {
                // spectra row y0 + rr holds image rows (y0 + rr - R, y0 + rr - R + D).  Rows past the last one (only the
                // discarded outputs yy >= D use them) fall into the next plane or the workspace behind Z: legal, unused.
                const float4 z = __ldg(reinterpret_cast<const float4 *>(zp + rr * zstride));  // (re0, re1, im0, im1)
                const float2 zre = make_float2(z.x, z.y), zim = make_float2(z.z, z.w);
#pragma unroll
                for (int y = 0; y < FT; ++y) {
                    const int dy = rr - R - y;
                    if (dy >= -R && dy <= R) {
                        are[y] = __ffma2_rn(kt[dy < 0 ? -dy : dy], zre, are[y]);
                        aim[y] = __ffma2_rn(kt[dy < 0 ? -dy : dy], zim, aim[y]);
                    }
                }
            }
            // filtered spectra stay split, 16 elements + 2 pad per block (16-byte aligned, conflict-free both ways)
            float4 *fo = reinterpret_cast<float4 *>(fism) + (p0 >> 1) + (p0 >> 4);
#pragma unroll
            for (int y = 0; y < FT; ++y) fo[y * (XB / 2)] = make_float4(are[y].x, are[y].y, aim[y].x, aim[y].y);
        }
        FFT_TRACE(3);
        if (i == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");  // this thread's twiddle copies (and staged state) have landed
        __syncthreads();
        FFT_TRACE(4);
        // phase 2: inverse transforms, 4 rows per pass; a row belongs to 64 threads (2 warps) with their own named barrier
        const float mu = __ldg(a.mu + pidx), sg = __ldg(a.sigma + pidx), wt = __ldg(a.weights + pidx);
        const float c2 = 1.4426950408889634f / (2.f * sg * sg);
#pragma unroll
        for (int ps = 0; ps < NP; ++ps) {
            float2 *buf = fism + (ps * RG + rw) * XB;
            const int yy = y0 + ps * RG + rw;
            float2 v[16];
            {
                // first DIT stage (span 1, no twiddles) on the split layout: quad m = (A, Cq) = (a.re b.re a.im b.im), (c.., d..)
                const float4 *b4 = reinterpret_cast<const float4 *>(buf) + 9 * t;
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const float4 A = b4[2 * m], Cq = b4[2 * m + 1];
                    const float2 S0 = cadd(make_float2(A.x, A.y), make_float2(Cq.x, Cq.y));  // (t0.re, t2.re)
                    const float2 S1 = cadd(make_float2(A.z, A.w), make_float2(Cq.z, Cq.w));  // (t0.im, t2.im)
                    const float2 D0 = csub(make_float2(A.x, A.y), make_float2(Cq.x, Cq.y));  // (t1.re, bd.re)
                    const float2 D1 = csub(make_float2(A.z, A.w), make_float2(Cq.z, Cq.w));  // (t1.im, bd.im)
                    v[4 * m] = make_float2(S0.x + S0.y, S1.x + S1.y);
                    v[4 * m + 1] = make_float2(D0.x - D1.y, D1.x + D0.y);
                    v[4 * m + 2] = make_float2(S0.x - S0.y, S1.x - S1.y);
                    v[4 * m + 3] = make_float2(D0.x + D1.y, D1.x - D0.y);
                }
            }
            row_sync(rw);  // the row's split spectra are consumed before the buffer changes layout
#pragma unroll
            for (int e = 0; e < 16; ++e) buf[17 * t + e] = v[e];
            row_sync(rw);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = buf[68 * B4 + k3 + 4 * q + (q >> 2)];
            r16_dit<4, 16>(v, tw_s, k3, k3, 4);
            row_sync(rw);
#pragma unroll
            for (int q = 0; q < 16; ++q) buf[64 * B4 + k3 + 4 * (q ^ (B4 & 3))] = v[q];
            row_sync(rw);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = buf[(t ^ (4 * (q & 3))) + 64 * q];
            // the row buffer is free now: the state values of the Euler step are staged into it by cp.async, their
            // latency hidden behind the last two stages + growth
            const bool stage = !EARLY && lenia && i == a.NS - 1;
            const float2 *svb = EARLY ? st_s + (ps * RG + rw) * N : buf;
            if (stage) {
                row_sync(rw);
                inv16_stage_state(a, buf, plane_off, min(yy, a.g.D - 1), seg, t);
            }
            r16_dit<64, 256>(v, tw_s, t, t, 64);
            // growth + weighted source sum (lenia.py:423-426, :444-446): 2*exp(-(u-mu)^2/(2 sigma^2)) - 1
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float d0 = v[q].x - mu, d1 = v[q].y - mu;
                const float g0 = 2.f * exp2f(-d0 * d0 * c2) - 1.f, g1 = 2.f * exp2f(-d1 * d1 * c2) - 1.f;
                if (MULTI) {
                    dx[ps][q].x = fmaf(g0, wt, dx[ps][q].x);
                    dx[ps][q].y = fmaf(g1, wt, dx[ps][q].y);
                } else {
                    v[q] = make_float2(g0 * wt, g1 * wt);
                }
            }
            FFT_TRACE(5);
            if (stage || (EARLY && lenia)) cp_async_commit_wait_all();  // own slots only: no barrier needed before reading them back
            FFT_TRACE(6);
            if (!MULTI && yy < a.g.D) {
                const bool rows_edge = a.row_wrap && (yy < F || yy >= a.g.D - F);
                const bool edge = rows_edge || a.g.periodic || seg == 0 || seg == a.g.nseg - 1;
                if (edge) inv16_store<true>(a, v, svb, plane_off, yy, seg, t);
                else inv16_store<false>(a, v, svb, plane_off, yy, seg, t);
            }
        }
    }
    FFT_TRACE(7);
    if (MULTI) {
#pragma unroll
        for (int ps = 0; ps < NP; ++ps) {
            const int yy = y0 + ps * RG + rw;
            if (yy >= a.g.D) continue;
            const float2 *buf = EARLY ? st_s + (ps * RG + rw) * N : fism + (ps * RG + rw) * XB;
            const bool rows_edge = a.row_wrap && (yy < F || yy >= a.g.D - F);
            const bool edge = rows_edge || a.g.periodic || seg == 0 || seg == a.g.nseg - 1;
            if (edge) inv16_store<true>(a, dx[ps], buf, plane_off, yy, seg, t);
            else inv16_store<false>(a, dx[ps], buf, plane_off, yy, seg, t);
        }
    }
}

// ---- plan helpers ---------------------------------------------------------------------------------
__global__ void lenia_fft_twiddle_kernel(float2 *tws, int N) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N - 1) return;
    // entry e lies in the stage with span s where (s-1) <= e < (4s-1)
    int span = 1;
    while (e >= 4 * span - 1) span <<= 2;
    const int off = e - (span - 1);
    const int r = off / span, k = off % span;
    double sn, cs;
    sincospi(-2.0 * (double)((r + 1) * k) / (double)(4 * span), &sn, &cs);
    tws[e] = make_float2((float)cs, (float)sn);
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
    if (W == 256 || W == 1024) {
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
    g.zr_begin = 0;
    g.zr_count = g.zrows;
    g.zr_gap_at = g.zrows;
    g.zr_gap = 0;
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

#ifdef B200CA_FFT_TRACE
extern "C" int b200ca_debug_fft_trace(long long *out, int n) {
    return (int)cudaMemcpyFromSymbol(out, g_fft_trace, sizeof(long long) * n);
}
#endif

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
    return b200ca_lenia_step_fft_phase(state_in, state_out, Kfft, mu, sigma, weights, workspace, B, C, k_mult, H, W, dt, mode, row_wrap,
                                       0, stream);
}

extern "C" int b200ca_lenia_step_fft_phase(const float *state_in, float *state_out, const float *Kfft, const float *mu,
                                           const float *sigma, const float *weights, float *workspace, int B, int C, int k_mult, int H,
                                           int W, float dt, int mode, int row_wrap, int phase, void *stream) {
    B200CA_REQUIRE(phase >= 0 && phase <= 2, "lenia_step_fft_phase: phase must be 0 (all), 1 (interior row FFTs) or 2 (rest)");
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
    // developer switches (timing / A-B runs), compiled in only with -DB200CA_DEV: B200CA_FFT_ONLY = 1|2|3 runs one kernel of
    // the step only; B200CA_FFT_FUSED = 0 unfuses FIR and inverse (N = 1024), 4|8 forces the rows per CTA of the fused kernel
    int rpc_rows_env = 2, rpc_inv = 1, only = 0, fused = -1, minb = 2, nt512 = 1;
    bool rpc_forced = false;
#ifdef B200CA_DEV
    {
        auto env = [](const char *n, int dflt) {
            const char *e = getenv(n);
            return e ? atoi(e) : dflt;
        };
        only = env("B200CA_FFT_ONLY", 0);
        fused = env("B200CA_FFT_FUSED", -1);
        rpc_inv = env("B200CA_FFT_RPC_INV", 1);
        if (rpc_inv != 2 && rpc_inv != 4) rpc_inv = 1;
        rpc_rows_env = env("B200CA_FFT_RPC_ROWS", 2);
        if (rpc_rows_env != 1 && rpc_rows_env != 4) rpc_rows_env = 2;
        rpc_forced = getenv("B200CA_FFT_RPC_ROWS") != nullptr;
        minb = env("B200CA_FFT_MINB", 2);
        nt512 = env("B200CA_FFT_NT512", 1);
    }
#endif
    if (only == 0 || only == 1) {
        // phase 1: spectra rows made of owned rows only (zr in [R, D+R)); phase 2: the 2R rows that touch ghost rows
        FftGeom gr = g;
        if (phase == 1) {
            gr.zr_begin = R;
            gr.zr_count = g.D;
        } else if (phase == 2) {
            gr.zr_begin = 0;
            gr.zr_count = 2 * R;
            gr.zr_gap_at = R;
            gr.zr_gap = g.D;
        }
#define ROWS4(RPC, NN, SP) launch_pdl(lenia_fft_rows_kernel<RPC, NN, SP>, dim3(ceil_div(gr.zr_count, RPC), g.nseg, B * C), dim3(NN / 4), \
                                  (size_t)RPC * NN * sizeof(float2), st, state_in, Z, tw, gr, H, W, pitch)
        const bool split = g.N == 1024 && fused != 0;  // the fused FIR + inverse kernel reads split spectra
        // small worlds (every CTA resident at once): one row pair per CTA shortens the serial chain (1024^2: 19.9 -> 19.4 us/step)
        const int rpc_rows = (!rpc_forced && (long long)gr.zr_count * g.nseg * B * C <= 5LL * num_sms()) ? 1 : rpc_rows_env;
        if (split) B200CA_CUDA(rpc_rows == 1 ? ROWS4(1, 1024, true) : (rpc_rows == 4 ? ROWS4(4, 1024, true) : ROWS4(2, 1024, true)));
        else if (g.N == 1024) B200CA_CUDA(rpc_rows == 1 ? ROWS4(1, 1024, false) : (rpc_rows == 4 ? ROWS4(4, 1024, false) : ROWS4(2, 1024, false)));
        else B200CA_CUDA(rpc_rows == 1 ? ROWS4(1, 256, false) : (rpc_rows == 4 ? ROWS4(4, 256, false) : ROWS4(2, 256, false)));
#undef ROWS4
        B200CA_LAUNCH_CHECK("lenia_fft_rows_kernel");
        note_variant("lenia_fft_rows<%d,%d,%s>", rpc_rows, g.N, split ? "split" : "plain");
    }
    if (phase == 1) return 0;
    FftInvArgs a;
    a.state_in = state_in;
    a.out = state_out;
    a.O = O;
    a.Z = Z;
    a.kfft = Kfft;
    a.k_mult = k_mult;
    a.tws = tw;
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
    if (g.N == 1024 && fused != 0) {
        if (only != 0 && only != 3) return 0;
        // 8 row pairs per CTA when that still gives every SM its resident CTAs, else 4 (more, smaller CTAs)
        const int ft = fused == 4 || fused == 8 ? fused : ((long long)ceil_div(g.D, 8) * g.nseg * B * C >= 2LL * num_sms() ? 8 : 4);
        const size_t smem = ((size_t)ft * XB + 1024) * sizeof(float2);  // row buffers + twiddle tables
        // (the opt-in shared-memory size is a per-device function attribute)
        static bool attr_done[64] = {false};
        int dev = 0;
        B200CA_CUDA(cudaGetDevice(&dev));
        if (!attr_done[dev & 63]) {
            const int big = (int)((8 * XB + 1024) * sizeof(float2));
            const int big5 = (int)(((size_t)8 * XB + 1024 + 8 * 1024) * sizeof(float2));
            B200CA_CUDA(cudaFuncSetAttribute(lenia_fft_firinv_kernel<8, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            B200CA_CUDA(cudaFuncSetAttribute(lenia_fft_firinv_kernel<8, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            B200CA_CUDA(cudaFuncSetAttribute(lenia_fft_firinv_kernel<8, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            B200CA_CUDA(cudaFuncSetAttribute(lenia_fft_firinv_kernel<8, false, 1, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, big5));
            B200CA_CUDA(cudaFuncSetAttribute(lenia_fft_firinv_kernel<8, true, 1, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, big5));
            attr_done[dev & 63] = true;
        }
        // small worlds: when every 8-row CTA fits the machine at one per SM the step is a single wave; 512-thread CTAs halve
        // the serial chain of a CTA (B200CA_FFT_NT512=0 disables)
        const long long ctas8 = (long long)ceil_div(g.D, 8) * g.nseg * B * C;
        if (nt512 && fused < 0 && ctas8 <= (long long)num_sms() * (nt512 >= 2 ? nt512 : 1)) {
            const size_t smem5 = ((size_t)8 * XB + 1024 + 8 * 1024) * sizeof(float2);  // row buffers + twiddles + staged state
            dim3 grid5(ceil_div(g.D, 8), g.nseg, B * C);
            B200CA_CUDA(NS > 1 ? launch_pdl(lenia_fft_firinv_kernel<8, true, 1, 512>, grid5, dim3(512), smem5, st, a)
                               : launch_pdl(lenia_fft_firinv_kernel<8, false, 1, 512>, grid5, dim3(512), smem5, st, a));
            B200CA_LAUNCH_CHECK("lenia_fft_firinv_kernel");
            note_variant("lenia_fft_firinv<8,%s,1,512>", NS > 1 ? "multi" : "single");
            return 0;
        }
        dim3 grid(ceil_div(g.D, ft), g.nseg, B * C);
        if (NS > 1) B200CA_CUDA(ft == 8 ? launch_pdl(lenia_fft_firinv_kernel<8, true, 2>, grid, dim3(256), smem, st, a)
                                             : launch_pdl(lenia_fft_firinv_kernel<4, true, 2>, grid, dim3(256), smem, st, a));
        else if (minb != 3) B200CA_CUDA(ft == 8 ? launch_pdl(lenia_fft_firinv_kernel<8, false, 2>, grid, dim3(256), smem, st, a)
                                                : launch_pdl(lenia_fft_firinv_kernel<4, false, 2>, grid, dim3(256), smem, st, a));
        else B200CA_CUDA(ft == 8 ? launch_pdl(lenia_fft_firinv_kernel<8, false, 3>, grid, dim3(256), smem, st, a)
                                 : launch_pdl(lenia_fft_firinv_kernel<4, false, 3>, grid, dim3(256), smem, st, a));
        B200CA_LAUNCH_CHECK("lenia_fft_firinv_kernel");
        note_variant("lenia_fft_firinv<%d,%s,%d,256>", ft, NS > 1 ? "multi" : "single", NS > 1 ? 2 : minb);
        return 0;
    }
    if (only == 0 || only == 2) {
        FftFirArgs f;
        f.Z = Z;
        f.O = O;
        f.kfft = Kfft;
        f.g = g;
        f.B = B;
        f.C = C;
        f.NS = NS;
        f.k_mult = k_mult;
        // 32 output rows per thread (1.9x read amplification) when that still fills the machine, else 8 (4.75x, from L2)
        const long long ctas32 = (long long)(g.N / FIR_THREADS) * g.nseg * ceil_div(g.D, 32) * B * NS * C;
        if (ctas32 >= (long long)num_sms() * 16) {
            dim3 grid((g.N / FIR_THREADS) * g.nseg, ceil_div(g.D, 32), B * NS * C);
            B200CA_CUDA(launch_pdl(lenia_fft_fir_kernel<32>, grid, dim3(FIR_THREADS), 0, st, f));
        } else {
            dim3 grid((g.N / FIR_THREADS) * g.nseg, ceil_div(g.D, 8), B * NS * C);
            B200CA_CUDA(launch_pdl(lenia_fft_fir_kernel<8>, grid, dim3(FIR_THREADS), 0, st, f));
        }
        B200CA_LAUNCH_CHECK("lenia_fft_fir_kernel");
        note_variant("lenia_fft_fir<%d>", ctas32 >= (long long)num_sms() * 16 ? 32 : 8);
    }
    if (only == 0 || only == 3) {
#define INV4(RPC, NN) launch_pdl(lenia_fft_inv_kernel<RPC, NN>, dim3(ceil_div(g.D, RPC), g.nseg, B * C), dim3(NN / 4), \
                                 (size_t)RPC * NN * sizeof(float2), st, a)
        if (g.N == 1024) B200CA_CUDA(rpc_inv == 1 ? INV4(1, 1024) : (rpc_inv == 4 ? INV4(4, 1024) : INV4(2, 1024)));
        else B200CA_CUDA(rpc_inv == 1 ? INV4(1, 256) : (rpc_inv == 4 ? INV4(4, 256) : INV4(2, 256)));
#undef INV4
        B200CA_LAUNCH_CHECK("lenia_fft_inv_kernel");
        note_variant("lenia_fft_inv<%d,%d>", rpc_inv, g.N);
    }
    return 0;
}
