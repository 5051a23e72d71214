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
#include "tma.cuh"

namespace b200ca {

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
};

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

template <int NW>
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

#pragma unroll 1
        for (int dy = 0; dy <= R; ++dy) {
            float sv[NLD * 4];
            {
                const float4 *rp = reinterpret_cast<const float4 *>(tl + dy * SROW);
#pragma unroll
                for (int q = 0; q < NLD; ++q) {
                    float4 v = rp[q];
                    sv[4 * q] = v.x; sv[4 * q + 1] = v.y; sv[4 * q + 2] = v.z; sv[4 * q + 3] = v.w;
                }
            }
            if (dy > 0) {
                const float4 *rm = reinterpret_cast<const float4 *>(tl - dy * SROW);
#pragma unroll
                for (int q = 0; q < NLD; ++q) {
                    float4 v = rm[q];
                    sv[4 * q] += v.x; sv[4 * q + 1] += v.y; sv[4 * q + 2] += v.z; sv[4 * q + 3] += v.w;
                }
            }
            float tv[KPITCH];
            {
                const float4 *kp = reinterpret_cast<const float4 *>(tp + dy * KPITCH);
#pragma unroll
                for (int q = 0; q < KPITCH / 4; ++q) {
                    float4 v = kp[q];
                    tv[4 * q] = v.x; tv[4 * q + 1] = v.y; tv[4 * q + 2] = v.z; tv[4 * q + 3] = v.w;
                }
            }
            float p[STRIP];
#pragma unroll
            for (int k = 0; k < STRIP; ++k) p[k] = tv[0] * sv[k + XOFF];
#pragma unroll
            for (int dx = 1; dx < KW; ++dx) {
#pragma unroll
                for (int k = 0; k < STRIP; ++k) p[k] = fmaf(tv[dx], sv[k + dx + XOFF], p[k]);
            }
#pragma unroll
            for (int k = 0; k < STRIP; ++k) acc[k] += p[k];
        }

        // growth + weighted source sum (lenia.py:423-426, :444-446)
        const int pidx = (b * a.NS + i) * a.C + j;
        const float mu = __ldg(a.mu + pidx), sg = __ldg(a.sigma + pidx), wt = __ldg(a.weights + pidx);
        const float sg2 = sg * sg;
#pragma unroll
        for (int k = 0; k < STRIP; ++k) {
            float d = acc[k] - mu;
            float z = __fdiv_rn(d * d, sg2) * 0.5f;
            float g = 2.f * expf(-z) - 1.f;
            dxs[k] += g * wt;
        }
        __syncthreads();  // everyone is done with tile[s] / taps[s] before they are refilled
    }

    // epilogue
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

template <int NW>
static int launch_conv(const CUtensorMap *tm, const LeniaArgs &a, int tiles_y, cudaStream_t st) {
    constexpr int TH = 8 * NW;
    constexpr size_t smem = (size_t)2 * (((TH + 2 * R) * SROW + 31) / 32 * 32) * 4 + 2 * KPREP * 4 + 64;
    static bool attr_set[64] = {false};
    int dev = 0;
    B200CA_CUDA(cudaGetDevice(&dev));
    if (!attr_set[dev & 63]) {
        B200CA_CUDA(cudaFuncSetAttribute(lenia_conv_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev & 63] = true;
    }
    dim3 grid(ceil_div(a.W, TW), tiles_y, a.B * a.C);
    lenia_conv_kernel<NW><<<grid, 32 * NW, smem, st>>>(tm, a);
    B200CA_LAUNCH_CHECK("lenia_conv_kernel");
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
    return 0;
}

extern "C" int b200ca_lenia_tile_rows(int H, int W, int B, int C) { return 8 * choose_nw(H, W, B, C); }

extern "C" int b200ca_lenia_step(const float *state_in, float *state_out, const float *Kprep, const float *mu, const float *sigma,
                                 const float *weights, int B, int C, int k_mult, int H, int W, float dt, int mode, int row_wrap,
                                 int tile_row_begin, int tile_row_end, void *stream) {
    B200CA_REQUIRE(state_in && state_out && Kprep && mu && sigma && weights, "lenia_step: null pointer");
    B200CA_REQUIRE(state_in != state_out, "lenia_step: in-place update is not supported");
    B200CA_REQUIRE(B >= 1 && C >= 1 && k_mult >= 1 && H >= F && W >= F, "lenia_step: need H,W >= %d (got %dx%d)", F, H, W);
    B200CA_REQUIRE((long long)B * C <= 65535, "lenia_step: B*C too large");
    B200CA_REQUIRE(mode == B200CA_MODE_LENIA || mode == B200CA_MODE_AFF, "lenia_step: bad mode %d", mode);
    const int nw = choose_nw(H, W, B, C);
    const int TH = 8 * nw;
    const int tiles_y_all = ceil_div(H, TH);
    if (tile_row_begin == 0 && tile_row_end == 0) tile_row_end = tiles_y_all;
    B200CA_REQUIRE(tile_row_begin >= 0 && tile_row_end <= tiles_y_all && tile_row_begin <= tile_row_end,
                   "lenia_step: tile row range [%d,%d) outside [0,%d)", tile_row_begin, tile_row_end, tiles_y_all);
    if (tile_row_begin == tile_row_end) return 0;
    LeniaArgs a;
    a.state_in = state_in;
    a.out = state_out;
    a.kprep = Kprep;
    a.mu = mu;
    a.sigma = sigma;
    a.weights = weights;
    a.B = B;
    a.C = C;
    a.NS = C * k_mult;
    a.k_mult = k_mult;
    a.H = H;
    a.W = W;
    a.pitch = b200ca_frame_pitch(W);
    a.dt = dt;
    a.mode = mode;
    a.row_wrap = row_wrap;
    a.tile_row0 = tile_row_begin;
    const CUtensorMap *tm = nullptr;
    const uint64_t rows = (uint64_t)H + 2 * F;
    int rc = get_tensor_map_3d(&tm, state_in, (uint64_t)a.pitch, rows, (uint64_t)B * C, (uint64_t)a.pitch, (uint64_t)a.pitch * rows,
                               SROW, (uint32_t)(TH + 2 * R));
    if (rc) return rc;
    const int ty = tile_row_end - tile_row_begin;
    cudaStream_t st = as_stream(stream);
    switch (nw) {
        case 2: return launch_conv<2>(tm, a, ty, st);
        case 4: return launch_conv<4>(tm, a, ty, st);
        default: return launch_conv<8>(tm, a, ty, st);
    }
}
