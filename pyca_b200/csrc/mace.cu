// MaCE-Lenia mass redistribution (+ cross-channel mixing) for sm_100a.
//
// Replaces the tail of MaCELenia._mace_step (pyca/automata/models/lenia/macelenia.py:103-118) and
// MaCELeniaXChan._cross_chan_step (lenia/mace_lenia_xchan.py:68-76).  The reference materialises
// two 9x im2col tensors (F.pad + F.unfold) per step; here one HBM-bound kernel reads the affinity
// and the state once (framed layouts: the periodic wrap is a plain read), keeps E = exp(beta*Aff)
// (halo 2) and A/Z (halo 1) of a 64x16 tile in shared memory, and writes the new state with its
// periodic frame.  Algorithmic traffic: Aff read + state read + state write = 12 B/cell/channel.
#include "common.cuh"

namespace b200ca {

constexpr int F = B200CA_FRAME;
constexpr int MTX = 64, MTY = 16, MTHREADS = 256;
constexpr int EW = MTX + 4, EH = MTY + 4;  // E tile (halo 2)
constexpr int GW = MTX + 2, GH = MTY + 2;  // A/Z tile (halo 1)

struct MaceArgs {
    const float *state_in, *aff;
    float *out;
    int B, C, H, W, pitch;
    float beta, alpha;
    int alpha_on, row_wrap;
};

__device__ __forceinline__ void mace_store(float *__restrict__ plane, int H, int W, int pitch, int y, int x, float v, bool row_wrap) {
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

__global__ void __launch_bounds__(MTHREADS) mace_kernel(const MaceArgs a) {
    extern __shared__ float sm[];
    float *E = sm;                 // [EH][EW]
    float *G = E + EH * EW;        // [GH][GW]
    float *outb = G + GH * GW;     // [C][MTY][MTX]   (only when alpha_on)
    float *affb = outb + (a.alpha_on ? a.C * MTY * MTX : 0);  // [C][MTY][MTX]
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * MTX, y0 = blockIdx.y * MTY, b = blockIdx.z;
    const int frows = a.H + 2 * F;
    const int lx = tid & (MTX - 1), ly0 = tid >> 6;  // 4 rows of threads, each handles rows ly0 + 4r

    for (int c = 0; c < a.C; ++c) {
        const size_t poff = (size_t)(b * a.C + c) * frows * a.pitch;
        const float *affp = a.aff + poff;
        const float *stp = a.state_in + poff;
        for (int t = tid; t < EH * EW; t += MTHREADS) {
            const int r = t / EW, q = t % EW;
            const int fy = y0 - 2 + r + F, fx = x0 - 2 + q + F;
            float v = 0.f;
            if (fy < frows && fx < a.pitch) v = expf(a.beta * __ldg(affp + (size_t)fy * a.pitch + fx));
            E[t] = v;
        }
        __syncthreads();
        for (int t = tid; t < GH * GW; t += MTHREADS) {
            const int r = t / GW, q = t % GW;
            const float *e = E + r * EW + q;  // top-left of the 3x3 window around (r+1, q+1) in E coords
            float z = e[0];
            z += e[1]; z += e[2];
            z += e[EW]; z += e[EW + 1]; z += e[EW + 2];
            z += e[2 * EW]; z += e[2 * EW + 1]; z += e[2 * EW + 2];
            const int fy = y0 - 1 + r + F, fx = x0 - 1 + q + F;
            float s = 0.f;
            if (fy < frows && fx < a.pitch) s = __ldg(stp + (size_t)fy * a.pitch + fx);
            G[t] = __fdiv_rn(s, z);
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < MTY / 4; ++rr) {
            const int ly = ly0 + 4 * rr;
            const float ec = E[(ly + 2) * EW + lx + 2];
            const float *g = G + ly * GW + lx;
            float acc = ec * g[0];
            acc += ec * g[1]; acc += ec * g[2];
            acc += ec * g[GW]; acc += ec * g[GW + 1]; acc += ec * g[GW + 2];
            acc += ec * g[2 * GW]; acc += ec * g[2 * GW + 1]; acc += ec * g[2 * GW + 2];
            const int y = y0 + ly, x = x0 + lx;
            if (a.alpha_on) {
                outb[(c * MTY + ly) * MTX + lx] = acc;
                float av = 0.f;
                if (y < a.H && x < a.W) av = __ldg(affp + (size_t)(y + F) * a.pitch + x + F);
                affb[(c * MTY + ly) * MTX + lx] = av;
            } else if (y < a.H && x < a.W) {
                mace_store(a.out + poff, a.H, a.W, a.pitch, y, x, acc, a.row_wrap != 0);
            }
        }
        __syncthreads();
    }
    if (!a.alpha_on) return;
    // cross-channel mixing: each thread owns its own cells in outb/affb (written by itself above)
#pragma unroll
    for (int rr = 0; rr < MTY / 4; ++rr) {
        const int ly = ly0 + 4 * rr;
        const int y = y0 + ly, x = x0 + lx;
        if (y >= a.H || x >= a.W) continue;
        float mx = -INFINITY, tot = 0.f;
        for (int c = 0; c < a.C; ++c) {
            mx = fmaxf(mx, affb[(c * MTY + ly) * MTX + lx]);
            tot += outb[(c * MTY + ly) * MTX + lx];
        }
        float den = 0.f;
        for (int c = 0; c < a.C; ++c) den += expf(a.beta * (affb[(c * MTY + ly) * MTX + lx] - mx));
        for (int c = 0; c < a.C; ++c) {
            const float num = expf(a.beta * (affb[(c * MTY + ly) * MTX + lx] - mx));
            const float p = __fdiv_rn(num, den);
            const float s = outb[(c * MTY + ly) * MTX + lx];
            const float target = tot * p;
            const float nv = s - (s - target) * a.alpha;
            const size_t poff = (size_t)(b * a.C + c) * frows * a.pitch;
            mace_store(a.out + poff, a.H, a.W, a.pitch, y, x, nv, a.row_wrap != 0);
        }
    }
}


// ---- fast path (C <= 4): register sliding windows + warp shuffles, no shared memory ---------------------------------
// A warp owns 28 output columns (lanes 2..29; lanes 0,1,30,31 carry the 2-cell halo) and marches down MROWS rows.  Per
// channel and input row r:  E(r) = exp(beta*Aff(r));  rE(r) = E(x-1)+E(x)+E(x+1) (2 shuffles);
// Z(r-1) = rE(r-2)+rE(r-1)+rE(r);  G(r-1) = A(r-1)/Z(r-1);  rG(r-1) (2 shuffles);
// out(r-2) = E(r-2) * (rG(r-3)+rG(r-2)+rG(r-1)).  ~40 instructions per cell and channel, every value loaded once.
constexpr int MCOLS = 28;

__device__ __noinline__ void mace_mirrors(float *__restrict__ plane, int H, int W, int pitch, int y, int x, float v, bool row_wrap) {
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

template <int C, int MROWS>
__global__ void __launch_bounds__(128) mace_fast_kernel(const MaceArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int xw = blockIdx.x * 4 + warp;
    const int x = xw * MCOLS - 2 + lane;           // column this lane carries (may be halo / outside)
    const int y0 = blockIdx.y * MROWS;
    const int b = blockIdx.z;
    const int frows = a.H + 2 * F;
    const int fx = min(max(x + F, 0), a.pitch - 1);  // clamped framed column (out-of-range lanes feed nothing valid)
    const bool out_lane = lane >= 2 && lane < 2 + MCOLS && x >= 0 && x < a.W;
    const float bl2 = a.beta * 1.4426950408889634f;
    const size_t pstride = (size_t)frows * a.pitch;
    const float *affp = a.aff + (size_t)b * C * pstride + fx;
    const float *stp = a.state_in + (size_t)b * C * pstride + fx;
    float *outp = a.out + (size_t)b * C * pstride;

    float E[C][3], rE[C][3], rG[C][3];  // index 0: oldest row
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int k = 0; k < 3; ++k) E[c][k] = rE[c][k] = rG[c][k] = 0.f;

    const int rend = min(y0 + MROWS, a.H) + 2;  // last input row + 1
    // software pipeline: the loads of input row r+1 are issued before row r is processed
    float afn[C], stn[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        afn[c] = __ldg(affp + (size_t)c * pstride + (size_t)(y0 - 2 + F) * a.pitch);
        stn[c] = __ldg(stp + (size_t)c * pstride + (size_t)(y0 - 3 + F) * a.pitch);
    }
    for (int r = y0 - 2; r < rend; ++r) {
        float afc[C], stc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            afc[c] = afn[c];
            stc[c] = stn[c];
        }
        {
            const int frn = min(r + 1 + F, frows - 1);  // next Aff row (clamped on the last iteration, value unused)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                afn[c] = __ldg(affp + (size_t)c * pstride + (size_t)frn * a.pitch);
                stn[c] = __ldg(stp + (size_t)c * pstride + (size_t)(frn - 1) * a.pitch);
            }
        }
        float outv[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float af = afc[c];   // Aff row r
            const float st = stc[c];   // state row r-1
            const float e = exp2f(bl2 * af);
            const float re = __shfl_up_sync(0xffffffffu, e, 1) + e + __shfl_down_sync(0xffffffffu, e, 1);
            // shift windows
            E[c][0] = E[c][1]; E[c][1] = E[c][2]; E[c][2] = e;
            rE[c][0] = rE[c][1]; rE[c][1] = rE[c][2]; rE[c][2] = re;
            const float z = rE[c][0] + rE[c][1] + rE[c][2];           // Z(r-1)
            const float gq = __fdividef(st, z);                        // G(r-1)
            const float rg = __shfl_up_sync(0xffffffffu, gq, 1) + gq + __shfl_down_sync(0xffffffffu, gq, 1);
            rG[c][0] = rG[c][1]; rG[c][1] = rG[c][2]; rG[c][2] = rg;
            outv[c] = E[c][0] * (rG[c][0] + rG[c][1] + rG[c][2]);      // out(r-2)
        }
        const int yo = r - 2;
        if (yo < y0 || !out_lane) continue;   // window still filling / halo lane
        if (a.alpha_on) {
            // softmax_c(beta * Aff_c) = E_c / sum_c E_c with the E = exp(beta * Aff) of row r-2 already in the window (the
            // reference subtracts the maximum first, mace_lenia_xchan.py:68-76; |beta * Aff| <= 8 needs no such guard in fp32)
            float tot = outv[0], den = E[0][0];
#pragma unroll
            for (int c = 1; c < C; ++c) {
                tot += outv[c];
                den += E[c][0];
            }
            const float inv = __fdividef(tot, den);
#pragma unroll
            for (int c = 0; c < C; ++c) outv[c] = outv[c] - (outv[c] - E[c][0] * inv) * a.alpha;
        }
        const bool edge = (x < F) || (x >= a.W - F) || (a.row_wrap && (yo < F || yo >= a.H - F));
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float *pl = outp + (size_t)c * pstride;
            pl[(size_t)(yo + F) * a.pitch + x + F] = outv[c];
            if (edge) mace_mirrors(pl, a.H, a.W, a.pitch, yo, x, outv[c], a.row_wrap != 0);
        }
    }
}

template <int C>
static int launch_mace_fast(const MaceArgs &a, cudaStream_t st) {
    // 16 rows per warp (25 % halo re-reads, all of them L2 hits) when that still gives every SM several blocks, else 8 rows.
    // (4096^2, C = 3, tools/mace_ab.py: 16 rows 177 us, 32 rows 184 us, 64 rows 212 us -- the shorter marches walk memory in a
    // tighter front)
    const long long blocks16 = (long long)ceil_div(ceil_div(a.W, MCOLS), 4) * ceil_div(a.H, 16) * a.B;
    if (blocks16 >= (long long)num_sms() * 16) {
        dim3 grid(ceil_div(ceil_div(a.W, MCOLS), 4), ceil_div(a.H, 16), a.B);
        mace_fast_kernel<C, 16><<<grid, 128, 0, st>>>(a);
        note_variant("mace_fast<%d,16>", C);
    } else {
        dim3 grid(ceil_div(ceil_div(a.W, MCOLS), 4), ceil_div(a.H, 8), a.B);
        mace_fast_kernel<C, 8><<<grid, 128, 0, st>>>(a);
        note_variant("mace_fast<%d,8>", C);
    }
    B200CA_LAUNCH_CHECK("mace_fast_kernel");
    return 0;
}

}  // namespace b200ca

using namespace b200ca;

extern "C" int b200ca_mace_redistribute(const float *state_in, const float *aff, float *state_out, int B, int C, int H, int W,
                                        float beta, int alpha_on, float alpha, int row_wrap, void *stream) {
    B200CA_REQUIRE(state_in && aff && state_out, "mace_redistribute: null pointer");
    B200CA_REQUIRE(state_in != state_out && aff != state_out, "mace_redistribute: output aliases an input");
    B200CA_REQUIRE(B >= 1 && C >= 1 && H >= F && W >= F, "mace_redistribute: need H,W >= %d", F);
    B200CA_REQUIRE(B <= 65535, "mace_redistribute: B too large");
    size_t smem = (size_t)(EH * EW + GH * GW + (alpha_on ? 2 * C * MTY * MTX : 0)) * sizeof(float);
    B200CA_REQUIRE(smem <= 200 * 1024, "mace_redistribute: too many channels (%d) for the cross-channel tile", C);
    if (smem > 48 * 1024) B200CA_CUDA(cudaFuncSetAttribute(mace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MaceArgs a;
    a.state_in = state_in;
    a.aff = aff;
    a.out = state_out;
    a.B = B;
    a.C = C;
    a.H = H;
    a.W = W;
    a.pitch = b200ca_frame_pitch(W);
    a.beta = beta;
    a.alpha = alpha;
    a.alpha_on = alpha_on ? 1 : 0;
    a.row_wrap = row_wrap;
    switch (C) {  // register/shuffle fast path for the channel counts PyCA ships; shared-memory kernel otherwise
        case 1: return launch_mace_fast<1>(a, as_stream(stream));
        case 2: return launch_mace_fast<2>(a, as_stream(stream));
        case 3: return launch_mace_fast<3>(a, as_stream(stream));
        case 4: return launch_mace_fast<4>(a, as_stream(stream));
        default: break;
    }
    dim3 grid(ceil_div(W, MTX), ceil_div(H, MTY), B);
    mace_kernel<<<grid, MTHREADS, smem, as_stream(stream)>>>(a);
    B200CA_LAUNCH_CHECK("mace_kernel");
    return 0;
}
