// MaCE-Lenia food subsystem (pyca/automata/models/lenia/macelenia.py:122-226) for sm_100a.
//
// Three elementwise kernels around the convolution / redistribution kernels of the MaCE step:
//   food_sense_kernel    Aff += (food > 0) + sum_src K_src,j (*) (food > 0)        (_compute_affinity with sense_food, :149-157)
//   food_decay_kernel    state -= (state > 0.02 ? 0.0003 : 0); per-world loss       (_decay_and_distribute, :171-192)
//   food_consume_kernel  transfer = min(food, rate * [food > 0 and sum_c state >= min_density]);
//                        state += transfer / 3; food -= transfer                     (_consume_food, :194-226, death off)
//                        + optionally the cross-channel mixing of MaCELeniaXChan, which the reference runs AFTER the
//                        food step (mace_lenia_xchan.py:59-76), so it cannot ride in the redistribution kernel here.
// The respawn of food spots when a world has lost 200 units of mass (:186-192) draws positions from Python's `random`
// on the host in the reference; it stays on the host (pyca_b200/lenia.py), between the decay and consume kernels.
// All of them are HBM-bound streaming passes (12-16 B per cell and channel); the state is framed, food is dense (B,1,H,W).
#include "common.cuh"

using namespace b200ca;

namespace {

constexpr int F = B200CA_FRAME;

// grid (ceil(pitch/256), H + 2F, B*C): the whole framed extent, so the frame of Aff stays consistent
__global__ void food_sense_kernel(float *__restrict__ aff, const float *__restrict__ fconv, const float *__restrict__ food, int C,
                                  int H, int W, int pitch) {
    const int fx = blockIdx.x * blockDim.x + threadIdx.x, fy = blockIdx.y, p = blockIdx.z;
    if (fx >= W + 2 * F) return;
    const int y = ((fy - F) % H + H) % H, x = ((fx - F) % W + W) % W;
    const size_t o = ((size_t)p * (H + 2 * F) + fy) * pitch + fx;
    const float m = food[((size_t)(p / C) * H + y) * W + x] > 0.f ? 1.f : 0.f;
    aff[o] = aff[o] + (m + fconv[o]);
}

// grid (ceil(W/256), H, B*C); interior only, in place (the consume kernel reads interior cells only and frames its output)
__global__ void food_decay_kernel(float *__restrict__ state, double *__restrict__ loss, int C, int H, int W, int pitch, float thresh,
                                  float decay) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, p = blockIdx.z;
    float d = 0.f;
    if (x < W) {
        const size_t o = ((size_t)p * (H + 2 * F) + y + F) * pitch + x + F;
        const float s = state[o];
        d = s > thresh ? decay : 0.f;
        state[o] = s - d;
    }
    // per-world loss: block sum -> one double atomic
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    __shared__ float ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += ws[w];
        if (t != 0.f) atomicAdd(loss + p / C, (double)t);
    }
}

// writes v at interior cell (y, x) of a framed plane and at its periodic images inside the frame
__device__ __forceinline__ void store_with_images(float *__restrict__ plane, int H, int W, int pitch, int y, int x, float v) {
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

// grid (ceil(W/256), H, B): one thread per interior cell, all channels; state_in -> state_out (different buffers, the
// output gets its periodic frame), food in place (each cell is read and written by its own thread only).
template <int MAXC>
__global__ void food_consume_kernel(const float *__restrict__ state_in, float *__restrict__ state_out, float *__restrict__ food,
                                    const float *__restrict__ aff, int C, int H, int W, int pitch, float min_density, float rate,
                                    int alpha_on, float beta, float alpha) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    const size_t plane = (size_t)(H + 2 * F) * pitch;
    const size_t src = (size_t)b * C * plane + (size_t)(y + F) * pitch + x + F;
    float s[MAXC];
    float tot = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
        if (c < C) {
            s[c] = state_in[src + c * plane];
            tot += s[c];
        }
    const size_t fo = ((size_t)b * H + y) * W + x;
    const float fd = food[fo];
    const bool overlap = fd > 0.f && tot >= min_density;
    const float transfer = fminf(fd, overlap ? rate : 0.f);
    const float add = transfer / 3.f;
    tot = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
        if (c < C) {
            s[c] += add;
            tot += s[c];
        }
    if (alpha_on) {
        // softmax_c(beta * Aff) of the PRE-update affinity; state <- state - (state - tot * P) * alpha (mace_lenia_xchan.py:68-76)
        float a[MAXC], mx = -3.4e38f, den = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) {
                a[c] = aff[src + c * plane];
                mx = fmaxf(mx, a[c]);
            }
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) {
                a[c] = expf(beta * (a[c] - mx));
                den += a[c];
            }
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) s[c] = s[c] - (s[c] - tot * __fdiv_rn(a[c], den)) * alpha;
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
        if (c < C) store_with_images(state_out + (size_t)(b * C + c) * plane, H, W, pitch, y, x, s[c]);
    food[fo] = fd - transfer;
}

}  // namespace

extern "C" int b200ca_mace_food_sense(float *aff, const float *food_conv, const float *food, int B, int C, int H, int W, void *stream) {
    B200CA_REQUIRE(aff && food_conv && food && B >= 1 && C >= 1 && H >= F && W >= F, "mace_food_sense: bad arguments");
    dim3 grid(ceil_div(W + 2 * F, 256), H + 2 * F, B * C);
    food_sense_kernel<<<grid, 256, 0, as_stream(stream)>>>(aff, food_conv, food, C, H, W, b200ca_frame_pitch(W));
    B200CA_LAUNCH_CHECK("food_sense_kernel");
    return 0;
}

extern "C" int b200ca_mace_food_decay(float *state, double *loss, int B, int C, int H, int W, float threshold, float decay, void *stream) {
    B200CA_REQUIRE(state && loss && B >= 1 && C >= 1 && H >= F && W >= F, "mace_food_decay: bad arguments");
    dim3 grid(ceil_div(W, 256), H, B * C);
    food_decay_kernel<<<grid, 256, 0, as_stream(stream)>>>(state, loss, C, H, W, b200ca_frame_pitch(W), threshold, decay);
    B200CA_LAUNCH_CHECK("food_decay_kernel");
    return 0;
}

extern "C" int b200ca_mace_food_consume(const float *state_in, float *state_out, float *food, const float *aff, int B, int C, int H,
                                        int W, float min_density, float transfer_rate, int alpha_on, float beta, float alpha,
                                        void *stream) {
    B200CA_REQUIRE(state_in && state_out && food && state_in != state_out, "mace_food_consume: bad pointers (out of place only)");
    B200CA_REQUIRE(!alpha_on || aff, "mace_food_consume: the cross-channel mixing needs the affinity");
    B200CA_REQUIRE(B >= 1 && C >= 1 && C <= 8 && H >= F && W >= F, "mace_food_consume: need 1 <= C <= 8, H,W >= %d", F);
    dim3 grid(ceil_div(W, 256), H, B);
    const int pitch = b200ca_frame_pitch(W);
    if (C <= 4)
        food_consume_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(state_in, state_out, food, aff, C, H, W, pitch, min_density,
                                                                    transfer_rate, alpha_on, beta, alpha);
    else
        food_consume_kernel<8><<<grid, 256, 0, as_stream(stream)>>>(state_in, state_out, food, aff, C, H, W, pitch, min_density,
                                                                    transfer_rate, alpha_on, beta, alpha);
    B200CA_LAUNCH_CHECK("food_consume_kernel");
    return 0;
}
