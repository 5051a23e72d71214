// Lenia-family step, direct (spatial) path, for sm_100a.
//
// Replaces Lenia.step / kernel_fftconv / growth (pyca/automata/models/lenia/lenia.py:430-469,
// :419-428) and MaCELenia._compute_affinity (lenia/macelenia.py:122-159).  The reference does the
// periodic 31x31 convolution with full-plane complex FFTs and ~14 elementwise passes over
// (B, C*k_mult, C, H, W) intermediates (~1-1.5 KB/cell of traffic).  Here ONE kernel per step
//   * TMA-loads a (TH+30) x 100 fp32 tile of each source channel from the framed state into shared
//     memory (double-buffered over the source channels, mbarrier completion);
//   * convolves it with the D4-symmetric kernel folded over rows: s = in[y+dy] + in[y-dy], so a
//     cell costs 16x31 FMAs instead of 31x31;  each thread owns a 1x16 output strip, reads its input
//     rows with conflict-free LDS.128 (row pitch 100 floats = 4 mod 32 banks) and the kernel taps
//     as warp-broadcast LDS.128;  per-row partial sums are added to the accumulator (two-level
//     summation keeps the fp32 error at the FFT path's level);
//   * applies the growth function, the weighted sum over source channels and the Euler step +
//     clamp (or writes the affinity for the MaCE family), with 128-bit stores, and maintains the
//     periodic frame of the output buffer so the next step needs no wrap logic.
// Algorithmic traffic 8 B/cell/channel; the kernel is bound by the fp32 FMA pipe (496 FMA per
// cell and channel pair), not by HBM -- see DESIGN.md.
#include <mutex>
#include <unordered_map>

#include "tma.cuh"

namespace b200ca {

// Kernel radius class of a prepared tap table, remembered per (device pointer) by b200ca_lenia_prepare_kernel so that the step
// entry point keeps its signature: the taps are embedded in 31 x 31 either way, the class only selects how much of that
// footprint the kernel evaluates (3, 5, 7, 9 or 15).  Unknown pointers (tables prepared elsewhere) take the full radius.
static std::mutex g_rk_mutex;
static std::unordered_map<const void *, int> g_rk_of;
static int radius_class(int ks) {
    const int r = ks / 2;
    return r <= 3 ? 3 : r <= 5 ? 5 : r <= 7 ? 7 : r <= 9 ? 9 : 15;
}
static int radius_of_table(const void *kprep) {
    std::lock_guard<std::mutex> lk(g_rk_mutex);
    auto it = g_rk_of.find(kprep);
    return it == g_rk_of.end() ? 15 : it->second;
}

constexpr int F = B200CA_FRAME;
constexpr int R = 15;            // kernel radius handled by the direct path (k_size <= 31)
constexpr int KW = 2 * R + 1;    // 31
constexpr int TW = 64;           // output tile width
constexpr int STRIP = 16;        // outputs per thread (one row)
constexpr int SROW = 100;        // smem row pitch in floats: >= TW + 2R, multiple of 4, == 4 (mod 32)
constexpr int NLD = 12;          // LDS.128 per input row segment (48 >= STRIP + 2R + 1 floats)
constexpr int XH = 16;           // left halo of the staged tile: the TMA box must start on a 16-byte boundary
                                 // (inner coordinate multiple of 4 floats; probed in tools/tma_probe.cu), so 16 not R
constexpr int XOFF = XH - R;     // column of the first tap inside a thread's 48-float row segment
constexpr int KROWS = R + 1;     // folded kernel rows dy = 0..15
constexpr int KPITCH = 32;       // floats per folded kernel row (31 taps + pad)
constexpr int KPREP = KROWS * KPITCH;

struct LeniaArgs {
    const float *state_in;
    float *out;
    const float *kprep, *mu, *sigma, *weights;
    int B, C, NS;  // NS = C * k_mult source kernels per destination
    int k_mult;
    int H, W, pitch;
    float dt;
    int mode, row_wrap;
    int tile_row0;  // first tile row of this launch
    // g_arbi growth (lenia.py:390-418, funcgen.py:79-129): a Fourier series of u per kernel pair instead of the Gaussian
    int gmode;                        // G_GAUSS, or (GARBI kernels) G_SERIES / G_RESCALE / G_STATS
    int nh;                           // harmonics per function
    const float *gcos, *gsin, *gfreq; // (B*NS*C, nh): cos / sin coefficients, fp32(2 pi / period) * harmonic
    float *gstats;                    // (B*NS*C, 2): min, max of the raw series over the plane (global rescale)
    float gr0, gspan;                 // rescale target: value * (r1 - r0) + r0
    float gclip;
    int gclip_on;
};
enum { G_GAUSS = 0, G_SERIES = 1, G_RESCALE = 2, G_STATS = 3, G_LINEAR = 4 };

__device__ __forceinline__ void atomic_min_f32(float *p, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int *>(p), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned *>(p), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f32(float *p, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int *>(p), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned *>(p), __float_as_uint(v));
}

// writes v at interior cell (y, x) of a framed plane and at its periodic images inside the frame
__device__ __forceinline__ void store_mirrors(float *__restrict__ plane, int H, int W, int pitch, int y, int x, float v,
                                              bool row_wrap, bool base_done) {
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
        for (int b = 0; b < nx; ++b) {
            if (base_done && a == 0 && b == 0) continue;
            plane[(size_t)ys[a] * pitch + xs[b]] = v;
        }
}

// GARBI = true: the growth of a pair is its Fourier series (g_arbi); kept out of the Gaussian instantiation, whose register
// allocation and code are what the direct path was tuned with.
// RK: radius of the kernel actually evaluated (k_size <= 2 RK + 1; the prepared taps are embedded in 31 x 31, zero outside): the
// folded rows dy = 0..RK and the taps dx = -RK..RK only -- (RK+1)(2RK+1) FMAs per cell and kernel pair instead of 16 x 31.
template <int NW, bool GARBI, int RK = R>
__global__ void __launch_bounds__(32 * NW) lenia_conv_kernel(const CUtensorMap *__restrict__ tmap, const LeniaArgs a) {
    constexpr int TH = 8 * NW;
    constexpr int TROWS = TH + 2 * R;
    constexpr int STAGE = (TROWS * SROW + 31) / 32 * 32;  // floats; stage buffers stay 128-byte aligned (TMA destination)
    constexpr int TX_BYTES = TROWS * SROW * 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *tile = reinterpret_cast<float *>(smem_raw);                 // [2][STAGE]
    float *taps = tile + 2 * STAGE;                                    // [2][KPREP]
    uint64_t *bars = reinterpret_cast<uint64_t *>(taps + 2 * KPREP);   // [2]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = lane & 7, tx = lane >> 3;
    const int x0 = blockIdx.x * TW;
    const int y0 = (blockIdx.y + a.tile_row0) * TH;
    const int b = blockIdx.z / a.C, j = blockIdx.z % a.C;
    const int yl = warp * 8 + ty;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bars[0], TX_BYTES);
        // framed coords of the tile origin: (x0 - XH + F, y0 - R + F); source plane of kernel 0 is channel 0
        tma_load_3d(tile, tmap, &bars[0], x0 + F - XH, y0 + F - R, b * a.C + 0);
    }

    float dxs[STRIP];
#pragma unroll
    for (int k = 0; k < STRIP; ++k) dxs[k] = 0.f;

    for (int i = 0; i < a.NS; ++i) {
        const int s = i & 1;
        if (tid == 0 && i + 1 < a.NS) {
            mbar_expect_tx(&bars[s ^ 1], TX_BYTES);
            tma_load_3d(tile + (s ^ 1) * STAGE, tmap, &bars[s ^ 1], x0 + F - XH, y0 + F - R, b * a.C + (i + 1) / a.k_mult);
        }
        // stage the folded taps of kernel (i -> j)
        {
            const float *src = a.kprep + ((size_t)(b * a.NS + i) * a.C + j) * KPREP;
            float *dst = taps + s * KPREP;
            for (int t = tid; t < KPREP / 4; t += 32 * NW)
                reinterpret_cast<float4 *>(dst)[t] = __ldg(reinterpret_cast<const float4 *>(src) + t);
        }
        mbar_wait(&bars[s], (i >> 1) & 1);
        __syncthreads();

        const float *tl = tile + s * STAGE + (yl + R) * SROW + tx * STRIP;
        const float *tp = taps + s * KPREP;
        float acc[STRIP];
#pragma unroll
        for (int k = 0; k < STRIP; ++k) acc[k] = 0.f;

        // columns of the row segment / taps that a radius-RK kernel touches (float4 granularity)
        constexpr int D0 = R - RK, D1 = R + RK;                                  // tap columns dx
        constexpr int Q0 = (D0 + XOFF) / 4, Q1 = (STRIP - 1 + D1 + XOFF) / 4;     // float4s of the input segment
        constexpr int T0 = D0 / 4, T1 = D1 / 4;                                   // float4s of the tap row
#pragma unroll 1
        for (int dy = 0; dy <= RK; ++dy) {
            float sv[NLD * 4];
            {
                const float4 *rp = reinterpret_cast<const float4 *>(tl + dy * SROW);
#pragma unroll
                for (int q = Q0; q <= Q1; ++q) {
                    float4 v = rp[q];
                    sv[4 * q] = v.x; sv[4 * q + 1] = v.y; sv[4 * q + 2] = v.z; sv[4 * q + 3] = v.w;
                }
            }
            if (dy > 0) {
                const float4 *rm = reinterpret_cast<const float4 *>(tl - dy * SROW);
#pragma unroll
                for (int q = Q0; q <= Q1; ++q) {
                    float4 v = rm[q];
                    sv[4 * q] += v.x; sv[4 * q + 1] += v.y; sv[4 * q + 2] += v.z; sv[4 * q + 3] += v.w;
                }
            }
            float tv[KPITCH];
            {
                const float4 *kp = reinterpret_cast<const float4 *>(tp + dy * KPITCH);
#pragma unroll
                for (int q = T0; q <= T1; ++q) {
                    float4 v = kp[q];
                    tv[4 * q] = v.x; tv[4 * q + 1] = v.y; tv[4 * q + 2] = v.z; tv[4 * q + 3] = v.w;
                }
            }
            float p[STRIP];
#pragma unroll
            for (int k = 0; k < STRIP; ++k) p[k] = tv[D0] * sv[k + D0 + XOFF];
#pragma unroll
            for (int dx = D0 + 1; dx <= D1; ++dx) {
#pragma unroll
                for (int k = 0; k < STRIP; ++k) p[k] = fmaf(tv[dx], sv[k + dx + XOFF], p[k]);
            }
#pragma unroll
            for (int k = 0; k < STRIP; ++k) acc[k] += p[k];
        }

        // growth + weighted source sum (lenia.py:423-426, :444-446)
        const int pidx = (b * a.NS + i) * a.C + j;
        if (!GARBI) {
            const float mu = __ldg(a.mu + pidx), sg = __ldg(a.sigma + pidx), wt = __ldg(a.weights + pidx);
            const float sg2 = sg * sg;
#pragma unroll
            for (int k = 0; k < STRIP; ++k) {
                float d = acc[k] - mu;
                float z = __fdiv_rn(d * d, sg2) * 0.5f;
                float g = 2.f * expf(-z) - 1.f;
                dxs[k] += g * wt;
            }
        } else {
            // values = sum_h cos_h cos(f_h u) + sin_h sin(f_h u);  optional (v - min) / (max - min + 1e-8) * (r1 - r0) + r0 with
            // the plane-wide min / max of the raw series (funcgen.py:99-117), optional clip from below
            if (a.gmode == G_LINEAR) {  // plain sum of the convolutions (food sensing, macelenia.py:152-154)
#pragma unroll
                for (int k = 0; k < STRIP; ++k) dxs[k] += acc[k];
                __syncthreads();
                continue;
            }
            const float wt = __ldg(a.weights + pidx);
            const float *gc = a.gcos + (size_t)pidx * a.nh, *gs = a.gsin + (size_t)pidx * a.nh, *gf = a.gfreq + (size_t)pidx * a.nh;
            float lo = 0.f, den = 1.f;
            if (a.gmode == G_RESCALE) {
                lo = a.gstats[2 * pidx];
                den = a.gstats[2 * pidx + 1] - lo + 1e-8f;
            }
            float vmin = __int_as_float(0x7f800000), vmax = __int_as_float(0xff800000);
            const bool row_ok = (y0 + yl) < a.H;
#pragma unroll 1
            for (int k = 0; k < STRIP; ++k) {
                const float u = acc[k];
                float v = 0.f;
                for (int h = 0; h < a.nh; ++h) {
                    float sn, cs;
                    sincosf(__ldg(gf + h) * u, &sn, &cs);
                    v += __ldg(gc + h) * cs + __ldg(gs + h) * sn;
                }
                if (a.gmode == G_STATS) {
                    if (row_ok && x0 + tx * STRIP + k < a.W) {
                        vmin = fminf(vmin, v);
                        vmax = fmaxf(vmax, v);
                    }
                    continue;
                }
                if (a.gmode == G_RESCALE) v = __fdiv_rn(v - lo, den) * a.gspan + a.gr0;
                if (a.gclip_on) v = fmaxf(v, a.gclip);
                dxs[k] += v * wt;
            }
            if (a.gmode == G_STATS) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
                    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
                }
                if (lane == 0 && vmin <= vmax) {
                    atomic_min_f32(a.gstats + 2 * pidx, vmin);
                    atomic_max_f32(a.gstats + 2 * pidx + 1, vmax);
                }
            }
        }
        __syncthreads();  // everyone is done with tile[s] / taps[s] before they are refilled
    }

    // epilogue
    if (GARBI && a.gmode == G_STATS) return;  // statistics pass: nothing is stored
    const int y = y0 + yl;
    if (y >= a.H) return;
    const int xb = x0 + tx * STRIP;
    if (xb >= a.W) return;
    const size_t plane_off = (size_t)(b * a.C + j) * (a.H + 2 * F) * a.pitch;
    float *oplane = a.out + plane_off;
    float res[STRIP];
    if (a.mode == B200CA_MODE_LENIA) {
        const float *srow = a.state_in + plane_off + (size_t)(y + F) * a.pitch + F + xb;
#pragma unroll
        for (int q = 0; q < STRIP / 4; ++q) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (xb + 4 * q + 3 < a.W + F) v = __ldg(reinterpret_cast<const float4 *>(srow) + q);  // frame keeps this in bounds
            float in4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float nv = in4[e] + a.dt * dxs[4 * q + e];
                res[4 * q + e] = fminf(fmaxf(nv, 0.f), 1.f);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < STRIP; ++k) res[k] = dxs[k];
    }
    float *orow = oplane + (size_t)(y + F) * a.pitch + F + xb;
    const bool full = (xb + STRIP <= a.W);
    if (full) {
#pragma unroll
        for (int q = 0; q < STRIP / 4; ++q)
            reinterpret_cast<float4 *>(orow)[q] = make_float4(res[4 * q], res[4 * q + 1], res[4 * q + 2], res[4 * q + 3]);
    }
    const bool edge = !full || (xb < F) || (xb + STRIP > a.W - F) || (a.row_wrap && (y < F || y >= a.H - F));
    if (edge) {
#pragma unroll
        for (int k = 0; k < STRIP; ++k)
            if (xb + k < a.W) store_mirrors(oplane, a.H, a.W, a.pitch, y, xb + k, res[k], a.row_wrap != 0, full);
    }
}

// ---- kernel preparation ------------------------------------------------------------------------
__global__ void lenia_prepare_kernel_kernel(const float *__restrict__ K, float *__restrict__ kprep, int nker, int ks,
                                            int *__restrict__ asym_flag) {
    const int ker = blockIdx.x;
    if (ker >= nker) return;
    const int off = (KW - ks) / 2;
    const float *k = K + (size_t)ker * ks * ks;
    float *o = kprep + (size_t)ker * KPREP;
    for (int t = threadIdx.x; t < KPREP; t += blockDim.x) {
        const int dy = t / KPITCH, c = t % KPITCH;
        float v = 0.f;
        if (c < KW) {
            const int rp = R + dy - off, rm = R - dy - off, cc = c - off;  // rows in the ks x ks kernel
            if (cc >= 0 && cc < ks && rp >= 0 && rp < ks && rm >= 0 && rm < ks) {
                const float up = k[rp * ks + cc], um = k[rm * ks + cc];
                v = 0.5f * (up + um);
                if (fabsf(up - um) > 1e-6f * fmaxf(fabsf(up), fabsf(um)) + 1e-12f) atomicExch(asym_flag, 1);
            }
        }
        o[t] = v;
    }
}

__global__ void garbi_stats_init_kernel(float *stats, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        stats[2 * i] = __int_as_float(0x7f800000);      // +inf
        stats[2 * i + 1] = __int_as_float(0xff800000);  // -inf
    }
}

template <int NW, bool GARBI, int RK = R>
static int launch_conv(const CUtensorMap *tm, const LeniaArgs &a, int tiles_y, cudaStream_t st) {
    constexpr int TH = 8 * NW;
    constexpr size_t smem = (size_t)2 * (((TH + 2 * R) * SROW + 31) / 32 * 32) * 4 + 2 * KPREP * 4 + 64;
    static bool attr_set[64] = {false};
    int dev = 0;
    B200CA_CUDA(cudaGetDevice(&dev));
    if (!attr_set[dev & 63]) {
        B200CA_CUDA(cudaFuncSetAttribute(lenia_conv_kernel<NW, GARBI, RK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev & 63] = true;
    }
    dim3 grid(ceil_div(a.W, TW), tiles_y, a.B * a.C);
    lenia_conv_kernel<NW, GARBI, RK><<<grid, 32 * NW, smem, st>>>(tm, a);
    B200CA_LAUNCH_CHECK("lenia_conv_kernel");
    note_variant("lenia_conv<%d,%s,R=%d>", NW, GARBI ? "garbi" : "gauss", RK);
    return 0;
}

// tile height choice: keep every SM equally loaded (the step is FMA-bound, so an uneven last wave is
// pure loss), prefer taller tiles (less halo re-staging) on ties.
static int choose_nw(int H, int W, int B, int C) {
    const int sms = num_sms();
    const int cand[3] = {8, 4, 2};
    int best = 8;
    double best_eff = -1.0;
    for (int c = 0; c < 3; ++c) {
        const int TH = 8 * cand[c];
        const long long tiles = (long long)ceil_div(W, TW) * ceil_div(H, TH) * B * C;
        const long long per_sm = ceil_div64(tiles, sms);
        double eff = (double)tiles / (double)(per_sm * sms);
        eff *= (double)H / (double)(ceil_div(H, TH) * TH);  // rows wasted in the last tile row
        if (eff > best_eff + 0.02) {
            best_eff = eff;
            best = cand[c];
        }
    }
    return best;
}

}  // namespace b200ca

using namespace b200ca;

extern "C" int64_t b200ca_lenia_kprep_elems(int B, int C, int k_mult) { return (int64_t)B * C * k_mult * C * KPREP; }

extern "C" int b200ca_lenia_prepare_kernel(const float *K, float *Kprep, int B, int C, int k_mult, int ks, void *stream) {
    B200CA_REQUIRE(K && Kprep, "lenia_prepare_kernel: null pointer");
    B200CA_REQUIRE(B >= 1 && C >= 1 && k_mult >= 1, "lenia_prepare_kernel: bad B/C/k_mult");
    if (ks < 1 || ks > KW || (ks % 2) == 0) {
        set_error("lenia_prepare_kernel: k_size %d unsupported by the direct path (odd, <= %d)", ks, KW);
        return B200CA_E_UNSUPPORTED;
    }
    cudaStream_t st = as_stream(stream);
    int *flag = nullptr;
    B200CA_CUDA(cudaMalloc(&flag, sizeof(int)));
    int rc = check_cuda(cudaMemsetAsync(flag, 0, sizeof(int), st), "memset");
    int h_flag = 0;
    if (!rc) {
        const int nker = B * C * k_mult * C;
        lenia_prepare_kernel_kernel<<<nker, 128, 0, st>>>(K, Kprep, nker, ks, flag);
        rc = check_cuda(cudaPeekAtLastError(), "lenia_prepare_kernel_kernel");
    }
    if (!rc) rc = check_cuda(cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st), "D2H flag");
    if (!rc) rc = check_cuda(cudaStreamSynchronize(st), "sync");
    cudaFree(flag);
    if (rc) return rc;
    if (h_flag) {
        set_error("lenia_prepare_kernel: kernel is not symmetric under row reversal; the folded direct path needs K[dy]==K[-dy]");
        return B200CA_E_UNSUPPORTED;
    }
    {
        std::lock_guard<std::mutex> lk(g_rk_mutex);
        g_rk_of[Kprep] = radius_class(ks);
    }
    return 0;
}

extern "C" int b200ca_lenia_tile_rows(int H, int W, int B, int C) { return 8 * choose_nw(H, W, B, C); }

static int lenia_step_impl(LeniaArgs a, int B, int C, int k_mult, int H, int W, int tile_row_begin, int tile_row_end,
                           void *stream) {
    B200CA_REQUIRE(a.state_in && a.out && a.kprep && a.weights, "lenia_step: null pointer");
    B200CA_REQUIRE(a.state_in != a.out, "lenia_step: in-place update is not supported");
    B200CA_REQUIRE(B >= 1 && C >= 1 && k_mult >= 1 && H >= F && W >= F, "lenia_step: need H,W >= %d (got %dx%d)", F, H, W);
    B200CA_REQUIRE((long long)B * C <= 65535, "lenia_step: B*C too large");
    B200CA_REQUIRE(a.mode == B200CA_MODE_LENIA || a.mode == B200CA_MODE_AFF, "lenia_step: bad mode %d", a.mode);
    const int nw = choose_nw(H, W, B, C);
    const int TH = 8 * nw;
    const int tiles_y_all = ceil_div(H, TH);
    if (tile_row_begin == 0 && tile_row_end == 0) tile_row_end = tiles_y_all;
    B200CA_REQUIRE(tile_row_begin >= 0 && tile_row_end <= tiles_y_all && tile_row_begin <= tile_row_end,
                   "lenia_step: tile row range [%d,%d) outside [0,%d)", tile_row_begin, tile_row_end, tiles_y_all);
    if (tile_row_begin == tile_row_end) return 0;
    a.B = B;
    a.C = C;
    a.NS = C * k_mult;
    a.k_mult = k_mult;
    a.H = H;
    a.W = W;
    a.pitch = b200ca_frame_pitch(W);
    a.tile_row0 = tile_row_begin;
    const CUtensorMap *tm = nullptr;
    const uint64_t rows = (uint64_t)H + 2 * F;
    int rc = get_tensor_map_3d(&tm, a.state_in, (uint64_t)a.pitch, rows, (uint64_t)B * C, (uint64_t)a.pitch, (uint64_t)a.pitch * rows,
                               SROW, (uint32_t)(TH + 2 * R));
    if (rc) return rc;
    const int ty = tile_row_end - tile_row_begin;
    cudaStream_t st = as_stream(stream);
    const int rk = radius_of_table(a.kprep);
    if (a.gmode == G_GAUSS) {
        switch (nw) {
#define B200CA_CONV_RK(NWV)                                                   \
    switch (rk) {                                                             \
        case 3: return launch_conv<NWV, false, 3>(tm, a, ty, st);             \
        case 5: return launch_conv<NWV, false, 5>(tm, a, ty, st);             \
        case 7: return launch_conv<NWV, false, 7>(tm, a, ty, st);             \
        case 9: return launch_conv<NWV, false, 9>(tm, a, ty, st);             \
        default: return launch_conv<NWV, false, R>(tm, a, ty, st);            \
    }
            case 2: B200CA_CONV_RK(2)
            case 4: B200CA_CONV_RK(4)
            default: B200CA_CONV_RK(8)
#undef B200CA_CONV_RK
        }
    }
    switch (nw) {
        case 2: return launch_conv<2, true>(tm, a, ty, st);
        case 4: return launch_conv<4, true>(tm, a, ty, st);
        default: return launch_conv<8, true>(tm, a, ty, st);
    }
}

extern "C" int b200ca_lenia_step(const float *state_in, float *state_out, const float *Kprep, const float *mu, const float *sigma,
                                 const float *weights, int B, int C, int k_mult, int H, int W, float dt, int mode, int row_wrap,
                                 int tile_row_begin, int tile_row_end, void *stream) {
    B200CA_REQUIRE(mu && sigma, "lenia_step: null pointer");
    LeniaArgs a = {};
    a.state_in = state_in;
    a.out = state_out;
    a.kprep = Kprep;
    a.mu = mu;
    a.sigma = sigma;
    a.weights = weights;
    a.dt = dt;
    a.mode = mode;
    a.row_wrap = row_wrap;
    a.gmode = G_GAUSS;
    return lenia_step_impl(a, B, C, k_mult, H, W, tile_row_begin, tile_row_end, stream);
}

extern "C" int b200ca_lenia_step_garbi(const float *state_in, float *state_out, const float *Kprep, const float *weights,
                                       const float *g_cos, const float *g_sin, const float *g_freq, int n_harmonics, int rescale_on,
                                       float rescale_lo, float rescale_hi, int clip_on, float clip_min, float *g_stats, int B, int C,
                                       int k_mult, int H, int W, float dt, int mode, int row_wrap, void *stream) {
    B200CA_REQUIRE(g_cos && g_sin && g_freq && n_harmonics >= 1 && n_harmonics <= 64, "lenia_step_garbi: bad growth series");
    B200CA_REQUIRE(!rescale_on || g_stats, "lenia_step_garbi: the global rescale needs the g_stats scratch (2 floats per kernel pair)");
    LeniaArgs a = {};
    a.state_in = state_in;
    a.out = state_out;
    a.kprep = Kprep;
    a.weights = weights;
    a.dt = dt;
    a.mode = mode;
    a.row_wrap = row_wrap;
    a.nh = n_harmonics;
    a.gcos = g_cos;
    a.gsin = g_sin;
    a.gfreq = g_freq;
    a.gstats = g_stats;
    a.gr0 = rescale_lo;
    a.gspan = rescale_hi - rescale_lo;
    a.gclip = clip_min;
    a.gclip_on = clip_on;
    if (rescale_on) {
        // pass 1: the same convolution, reduced to the min / max of every pair's raw series over its plane; pass 2 recomputes
        // the (bit-identical) values and rescales them
        const int npairs = B * C * k_mult * C;
        garbi_stats_init_kernel<<<ceil_div(npairs, 128), 128, 0, as_stream(stream)>>>(g_stats, npairs);
        B200CA_LAUNCH_CHECK("garbi_stats_init_kernel");
        a.gmode = G_STATS;
        const int rc = lenia_step_impl(a, B, C, k_mult, H, W, 0, 0, stream);
        if (rc) return rc;
    }
    a.gmode = rescale_on ? G_RESCALE : G_SERIES;
    return lenia_step_impl(a, B, C, k_mult, H, W, 0, 0, stream);
}

extern "C" int b200ca_lenia_conv_sum(const float *planes_in, float *out, const float *Kprep, int B, int C, int k_mult, int H, int W,
                                     void *stream) {
    LeniaArgs a = {};
    a.state_in = planes_in;
    a.out = out;
    a.kprep = Kprep;
    a.weights = Kprep;  // unused in this mode
    a.mode = B200CA_MODE_AFF;
    a.row_wrap = 1;
    a.gmode = G_LINEAR;
    return lenia_step_impl(a, B, C, k_mult, H, W, 0, 0, stream);
}
