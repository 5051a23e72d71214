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
    dim3 grid(ceil_div(W, MTX), ceil_div(H, MTY), B);
    mace_kernel<<<grid, MTHREADS, smem, as_stream(stream)>>>(a);
    B200CA_LAUNCH_CHECK("mace_kernel");
    return 0;
}
