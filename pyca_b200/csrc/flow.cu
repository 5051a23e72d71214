// FlowLenia step tail (pyca/automata/models/lenia/flow_lenia.py:97-153, :207-253) for sm_100a: flow field + reintegration
// tracking in one kernel.  The affinity comes from the Lenia convolution kernels (mode AFF), exactly as for MaCE.
//
// Reference, per step and channel: two zero-padded Sobel convolutions of Aff, two of the total mass, alpha, the flow
// F = grad(Aff) (1 - alpha) - grad(mass) alpha, then 25 rolled copies of (state, mu) with an overlap product each --
// ~60 full-tensor torch ops and ~100 temporaries.  Here one CTA owns a 32x16 output tile: it stages the total mass and the
// channel's Aff on the tile + halo (dd + 1), derives the clipped displacements clip(dt F) and the state on the tile + halo
// dd in shared memory, and every thread then gathers its (2dd+1)^2 sources from shared memory.  HBM traffic is the
// algorithmic one: state + Aff read (tile halos are L2 hits), state written.
//
// Two properties of the reference are kept on purpose: (1) the Sobel stencils are ZERO padded (conv2d padding="same"),
// (2) positions are absolute (x + 0.5), so a source rolled in across the world edge never overlaps its destination: the
// flow is not periodic even though the Lenia convolution in front of it is.  Position arithmetic follows the reference's
// order of fp32 operations (mu = pos + clip(dt F); |mu - pos'|): at x ~ 1000 one ulp of a position is 6e-5.
#include "common.cuh"

using namespace b200ca;

namespace {

constexpr int F = B200CA_FRAME;
constexpr int FTX = 32, FTY = 32, FTHREADS = 256, RPT = 4;  // 32x32 output tile, 4 vertically adjacent outputs per thread

struct FlowArgs {
    const float *state_in, *aff;
    float *out;
    int B, C, H, W, pitch;
    float dt, sigma, max_flow, cap, inv_norm;  // cap = min(2 sigma, 1); inv_norm = 1 / (4 sigma^2)
    float theta_x;
    int n_pow;
};

__device__ __forceinline__ void flow_store(float *__restrict__ plane, int H, int W, int pitch, int y, int x, float v) {
    const int fy = y + F, fx = x + F;
    int ys[3], xs[3], ny = 0, nx = 0;
    ys[ny++] = fy;
    if (y < F) ys[ny++] = fy + H;
    if (y >= H - F) ys[ny++] = fy - H;
    xs[nx++] = fx;
    if (x < F) xs[nx++] = fx + W;
    if (x >= W - F) xs[nx++] = fx - W;
    for (int a = 0; a < ny; ++a)
        for (int b = 0; b < nx; ++b) plane[(size_t)ys[a] * pitch + xs[b]] = v;
}

// Sobel cross-correlations of flow_lenia.py:12-29 on a shared-memory tile with row pitch `p`, centred at t
__device__ __forceinline__ void sobel_at(const float *t, int p, float &sw, float &sh) {
    const float a = t[-p - 1], b = t[-p], c = t[-p + 1], d = t[-1], e = t[1], f = t[p - 1], g = t[p], h = t[p + 1];
    sw = (c - a) + 2.f * (e - d) + (h - f);
    sh = (f - a) + 2.f * (g - b) + (h - c);
}

// UNIT_CAP: min(2 sigma, 1) == 1 (sigma >= 0.5, the default 0.65): the clip to [0, cap] is the free .SAT modifier
template <int DD, bool UNIT_CAP>
__global__ void __launch_bounds__(FTHREADS, 4) flow_rt_kernel(const FlowArgs a) {
    constexpr int HA = FTY + 2 * DD + 2, WA = FTX + 2 * DD + 2;  // Aff / mass tiles: halo DD + 1
    constexpr int HS = FTY + 2 * DD, WS = FTX + 2 * DD;          // displacement / state tiles: halo DD
    __shared__ float mass[HA * WA], affs[HA * WA];
    __shared__ float4 src[HS * WS];  // per source cell: (clipped displacement x, y, state, -)
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * FTX, y0 = blockIdx.y * FTY, b = blockIdx.z;
    const size_t plane = (size_t)(a.H + 2 * F) * a.pitch;
    const float *sb = a.state_in + (size_t)b * a.C * plane;
    const float *ab = a.aff + (size_t)b * a.C * plane;

    for (int t = tid; t < HA * WA; t += FTHREADS) {
        const int y = y0 - DD - 1 + t / WA, x = x0 - DD - 1 + t % WA;
        float m = 0.f;
        if (y >= 0 && y < a.H && x >= 0 && x < a.W) {
            const size_t o = (size_t)(y + F) * a.pitch + x + F;
            for (int c = 0; c < a.C; ++c) m += sb[o + c * plane];
        }
        mass[t] = m;
    }
    const int lx = tid & 31, ly0 = (tid >> 5) * RPT;
    for (int c = 0; c < a.C; ++c) {
        __syncthreads();  // previous channel's gathers are done with src / affs
        for (int t = tid; t < HA * WA; t += FTHREADS) {
            const int y = y0 - DD - 1 + t / WA, x = x0 - DD - 1 + t % WA;
            affs[t] = (y >= 0 && y < a.H && x >= 0 && x < a.W) ? ab[c * plane + (size_t)(y + F) * a.pitch + x + F] : 0.f;
        }
        __syncthreads();
        for (int t = tid; t < HS * WS; t += FTHREADS) {
            const int r = t / WS, q = t % WS;
            const int y = y0 - DD + r, x = x0 - DD + q;
            float sw, sh, gw, gh;
            sobel_at(affs + (r + 1) * WA + q + 1, WA, sw, sh);
            // grad(total mass) and alpha = clip((mass / theta_x)^n, 0, 1): recomputed per channel from the mass tile (a few
            // shared-memory reads) rather than kept in three more tiles
            const float *mc = mass + (r + 1) * WA + q + 1;
            sobel_at(mc, WA, gw, gh);
            const float mv = mc[0] / a.theta_x;
            const float al = fminf(fmaxf(a.n_pow == 2 ? mv * mv : powf(mv, (float)a.n_pow), 0.f), 1.f), om = 1.f - al;
            // F = grad_u * (1 - alpha) - grad_x * alpha, every product rounded on its own as torch does
            const float fx = __fsub_rn(__fmul_rn(sw, om), __fmul_rn(gw, al));
            const float fy = __fsub_rn(__fmul_rn(sh, om), __fmul_rn(gh, al));
            // a source outside the world is a rolled-in one in the reference: its absolute position never overlaps
            const float sv = (y >= 0 && y < a.H && x >= 0 && x < a.W) ? sb[c * plane + (size_t)(y + F) * a.pitch + x + F] : 0.f;
            src[t] = make_float4(fminf(fmaxf(__fmul_rn(a.dt, fx), -a.max_flow), a.max_flow),
                                 fminf(fmaxf(__fmul_rn(a.dt, fy), -a.max_flow), a.max_flow), sv, 0.f);
        }
        __syncthreads();
        // gather: a thread owns RPT vertically adjacent outputs of one column, so a source cell is read once for all the
        // outputs it feeds and its x-overlap (same column distance for all of them) is computed once
        {
            const int x = x0 + lx, yb = y0 + ly0;
            const float px = (float)x + 0.5f, pyb = (float)yb + 0.5f;  // cell centres: small offsets from them stay exact
            float acc[RPT];
#pragma unroll
            for (int k = 0; k < RPT; ++k) acc[k] = 0.f;
#pragma unroll
            for (int dx = -DD; dx <= DD; ++dx) {
#pragma unroll
                for (int sr = -DD; sr < RPT + DD; ++sr) {  // source row relative to the thread's first output row
                    const float4 s4 = src[(ly0 + sr + DD) * WS + (lx - dx + DD)];
                    // mu of the source = its own centre + its displacement, rounded as the reference rounds it (the centres
                    // are exact in fp32); overlap of its sigma-square with the destination's unit square, per axis
                    const float mux = __fadd_rn(px - (float)dx, s4.x);
                    float ox = __saturatef(0.5f + (a.sigma - fabsf(mux - px)));
                    if (!UNIT_CAP) ox = fminf(ox, a.cap);
                    const float muy = __fadd_rn(pyb + (float)sr, s4.y);
                    const float w = s4.z * ox;
#pragma unroll
                    for (int k = 0; k < RPT; ++k) {
                        if (k - sr < -DD || k - sr > DD) continue;  // dy = k - sr
                        float oy = __saturatef(0.5f + (a.sigma - fabsf(muy - (pyb + (float)k))));
                        if (!UNIT_CAP) oy = fminf(oy, a.cap);
                        acc[k] = fmaf(w, oy, acc[k]);
                    }
                }
            }
            if (x < a.W) {
#pragma unroll
                for (int k = 0; k < RPT; ++k)
                    if (yb + k < a.H)
                        flow_store(a.out + (size_t)(b * a.C + c) * plane, a.H, a.W, a.pitch, yb + k, x, acc[k] * a.inv_norm);
            }
        }
    }
}

template <int DD>
int launch_flow(const FlowArgs &a, cudaStream_t st) {
    dim3 grid(ceil_div(a.W, FTX), ceil_div(a.H, FTY), a.B);
    if (a.cap == 1.f) flow_rt_kernel<DD, true><<<grid, FTHREADS, 0, st>>>(a);
    else flow_rt_kernel<DD, false><<<grid, FTHREADS, 0, st>>>(a);
    B200CA_LAUNCH_CHECK("flow_rt_kernel");
    note_variant("flow_rt<%d,%s>", DD, a.cap == 1.f ? "unit" : "cap");
    return 0;
}

}  // namespace

extern "C" int b200ca_flow_reintegrate(const float *state_in, const float *aff, float *state_out, int B, int C, int H, int W, float dt,
                                       int dd, float sigma_rt, float theta_x, int n_pow, void *stream) {
    B200CA_REQUIRE(state_in && aff && state_out && state_in != state_out, "flow_reintegrate: bad pointers (out of place only)");
    B200CA_REQUIRE(B >= 1 && B <= 65535 && C >= 1 && H >= F && W >= F, "flow_reintegrate: need H,W >= %d", F);
    B200CA_REQUIRE(dd >= 1 && dd <= 3, "flow_reintegrate: dd must be 1, 2 or 3 (got %d)", dd);
    B200CA_REQUIRE(sigma_rt > 0.f && n_pow >= 1, "flow_reintegrate: bad sigma_rt / n");
    FlowArgs a;
    a.state_in = state_in;
    a.aff = aff;
    a.out = state_out;
    a.B = B, a.C = C, a.H = H, a.W = W;
    a.pitch = b200ca_frame_pitch(W);
    a.dt = dt;
    a.sigma = sigma_rt;
    // the reference forms these in Python doubles and torch casts them to fp32 when they meet the tensors
    a.max_flow = (float)((double)dd - (double)sigma_rt);
    a.cap = (float)(2.0 * (double)sigma_rt < 1.0 ? 2.0 * (double)sigma_rt : 1.0);
    a.inv_norm = 1.f / (float)(4.0 * (double)sigma_rt * (double)sigma_rt);
    a.theta_x = theta_x;
    a.n_pow = n_pow;
    cudaStream_t st = as_stream(stream);
    switch (dd) {
        case 1: return launch_flow<1>(a, st);
        case 2: return launch_flow<2>(a, st);
        default: return launch_flow<3>(a, st);
    }
}
