// Neural-CA inference step for sm_100a: weight preparation, the fp32 SIMT update kernel, the
// life-mask pass and the C entry points.  (The tcgen05/TMEM update kernel lives in nca_tc.cu.)
//
// Replaces NCAModule.forward for one step (pyca/automata/models/nca.py:239-263):
//   perception  SobelConv.forward            nca.py:328-344  (grouped conv quirk: output o reads channel o//2)
//   MLP         Linear(48,128)-ReLU-Linear(128,16)   nca.py:205-208
//   update      sigmoid(dx) * (rand <= 0.5)  nca.py:256-259
//   life mask   alive_before & alive_after, 3x3 circular max-pool of channel 3 > 0.1   nca.py:226-237, :260-263
// The reference runs ~20 ATen kernels over (B,48,H,W) / (B*H*W,128) intermediates.  Here the update pass
// reads the state once and writes it once; `alive_after` needs the NEW alpha of the 8 neighbours, so the
// update pass also emits two bit planes (new alpha > 0.1; cell written non-zero) and a second, tiny pass
// zeroes the cells that died (reads 2 bits/cell, writes only the dying cells).
#include "nca.cuh"

namespace b200ca {

constexpr int SIMT_THREADS = 128;
constexpr int SIMT_ROWS = 2;  // cells per thread (vertically adjacent)

__device__ __forceinline__ float sigmoidf_rn(float x) { return __fdiv_rn(1.f, 1.f + expf(-x)); }

__global__ void __launch_bounds__(SIMT_THREADS) nca_update_simt_kernel(const NcaArgs a) {
    __shared__ __align__(16) float w[WP_SIMT_FLOATS];
    for (int t = threadIdx.x; t < WP_SIMT_FLOATS / 4; t += SIMT_THREADS)
        reinterpret_cast<float4 *>(w)[t] = __ldg(reinterpret_cast<const float4 *>(a.wp) + t);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 32 + lane;
    const int ybase = (blockIdx.y * (SIMT_THREADS / 32) + warp) * SIMT_ROWS;
    const int b = blockIdx.z;
    const bool xin = x < a.W;
    const int xc = xin ? x : a.W - 1;
    const int xm = xc == 0 ? a.W - 1 : xc - 1, xp = xc == a.W - 1 ? 0 : xc + 1;
    const size_t plane = (size_t)a.H * a.W;
    const float *sb = a.in + (size_t)b * NCA_N * plane;

    float p[SIMT_ROWS][NCA_K1];
    bool alive0[SIMT_ROWS], rowin[SIMT_ROWS];
#pragma unroll
    for (int r = 0; r < SIMT_ROWS; ++r) {
        const int y = ybase + r;
        rowin[r] = y < a.H;
        const int yc = rowin[r] ? y : a.H - 1;
        const int ym = yc == 0 ? a.H - 1 : yc - 1, yp = yc == a.H - 1 ? 0 : yc + 1;
#pragma unroll
        for (int c = 0; c < NCA_N; ++c) {
            const float *pl = sb + (size_t)c * plane;
            const float ctr = __ldg(pl + (size_t)yc * a.W + xc);
            p[r][16 + c] = ctr;
            if (c < 8) {  // sobel-x / 8
                const float l0 = __ldg(pl + (size_t)ym * a.W + xm), r0 = __ldg(pl + (size_t)ym * a.W + xp);
                const float l1 = __ldg(pl + (size_t)yc * a.W + xm), r1 = __ldg(pl + (size_t)yc * a.W + xp);
                const float l2 = __ldg(pl + (size_t)yp * a.W + xm), r2 = __ldg(pl + (size_t)yp * a.W + xp);
                p[r][c] = ((r0 - l0) + 2.f * (r1 - l1) + (r2 - l2)) * 0.125f;
                if (c == 3) {
                    const float t1 = __ldg(pl + (size_t)ym * a.W + xc), b1 = __ldg(pl + (size_t)yp * a.W + xc);
                    float m = fmaxf(fmaxf(fmaxf(l0, r0), fmaxf(l1, r1)), fmaxf(fmaxf(l2, r2), fmaxf(t1, b1)));
                    alive0[r] = fmaxf(m, ctr) > 0.1f;
                }
            } else {  // sobel-y / 8
                const float t0 = __ldg(pl + (size_t)ym * a.W + xm), t1 = __ldg(pl + (size_t)ym * a.W + xc),
                            t2 = __ldg(pl + (size_t)ym * a.W + xp);
                const float b0 = __ldg(pl + (size_t)yp * a.W + xm), b1 = __ldg(pl + (size_t)yp * a.W + xc),
                            b2 = __ldg(pl + (size_t)yp * a.W + xp);
                p[r][c] = ((b0 - t0) + 2.f * (b1 - t1) + (b2 - t2)) * 0.125f;
            }
        }
    }

    float o[SIMT_ROWS][NCA_N];
#pragma unroll
    for (int r = 0; r < SIMT_ROWS; ++r)
#pragma unroll
        for (int c = 0; c < NCA_N; ++c) o[r][c] = w[WP_B2 + c];

#pragma unroll 2
    for (int k = 0; k < NCA_HID; ++k) {
        float wr[NCA_K1];
        const float4 *w1 = reinterpret_cast<const float4 *>(w + WP_W1F + k * NCA_K1);
#pragma unroll
        for (int q = 0; q < NCA_K1 / 4; ++q) {
            float4 v = w1[q];
            wr[4 * q] = v.x; wr[4 * q + 1] = v.y; wr[4 * q + 2] = v.z; wr[4 * q + 3] = v.w;
        }
        float h[SIMT_ROWS];
#pragma unroll
        for (int r = 0; r < SIMT_ROWS; ++r) {
            float s = w[WP_B1 + k];
#pragma unroll
            for (int f = 0; f < NCA_K1; ++f) s = fmaf(wr[f], p[r][f], s);
            h[r] = fmaxf(s, 0.f);
        }
        const float4 *w2 = reinterpret_cast<const float4 *>(w + WP_W2T + k * NCA_N);
#pragma unroll
        for (int q = 0; q < NCA_N / 4; ++q) {
            float4 v = w2[q];
#pragma unroll
            for (int r = 0; r < SIMT_ROWS; ++r) {
                o[r][4 * q] = fmaf(v.x, h[r], o[r][4 * q]);
                o[r][4 * q + 1] = fmaf(v.y, h[r], o[r][4 * q + 1]);
                o[r][4 * q + 2] = fmaf(v.z, h[r], o[r][4 * q + 2]);
                o[r][4 * q + 3] = fmaf(v.w, h[r], o[r][4 * q + 3]);
            }
        }
    }

    float *ob = a.out + (size_t)b * NCA_N * plane;
#pragma unroll
    for (int r = 0; r < SIMT_ROWS; ++r) {
        const int y = ybase + r;
        const bool valid = xin && rowin[r];
        const size_t cell = ((size_t)b * a.H + (rowin[r] ? y : 0)) * a.W + xc;
        const bool upd = valid && nca_update_bit(a, cell);
        const float alpha_new = upd ? sigmoidf_rn(o[r][3]) : 0.f;
        const bool keep = upd && alive0[r];
        const uint32_t abits = __ballot_sync(0xffffffffu, valid && alpha_new > 0.1f);
        const uint32_t nbits = __ballot_sync(0xffffffffu, keep);
        if (rowin[r] && lane == 0) {
            const size_t wi = ((size_t)b * a.H + y) * a.wpr + blockIdx.x;
            a.alive_bits[wi] = abits;
            a.nz_bits[wi] = nbits;
        }
        if (valid) {
#pragma unroll
            for (int c = 0; c < NCA_N; ++c) ob[(size_t)c * plane + (size_t)y * a.W + x] = keep ? sigmoidf_rn(o[r][c]) : 0.f;
        }
    }
}

// zeroes the cells that were written non-zero but have no alive (new alpha > 0.1) cell in their 3x3 torus window
__global__ void nca_life_kernel(float *__restrict__ state, const uint32_t *__restrict__ alive_bits,
                                const uint32_t *__restrict__ nz_bits, int B, int H, int W, int wpr) {
    pdl_wait_then_release();   // the bit planes and the state are the update kernel's output
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    const size_t rowbase = (size_t)b * H;
    if (!((nz_bits[(rowbase + y) * wpr + (x >> 5)] >> (x & 31)) & 1u)) return;
    const int ym = y == 0 ? H - 1 : y - 1, yp = y == H - 1 ? 0 : y + 1;
    const int xm = x == 0 ? W - 1 : x - 1, xp = x == W - 1 ? 0 : x + 1;
    const int ys[3] = {ym, y, yp}, xs[3] = {xm, x, xp};
    bool alive = false;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            alive |= ((__ldg(alive_bits + (rowbase + ys[i]) * wpr + (xs[j] >> 5)) >> (xs[j] & 31)) & 1u) != 0;
    if (alive) return;
    const size_t plane = (size_t)H * W;
    float *p = state + (size_t)b * NCA_N * plane + (size_t)y * W + x;
#pragma unroll
    for (int c = 0; c < NCA_N; ++c) p[(size_t)c * plane] = 0.f;
}

__global__ void nca_prepare_simt_kernel(const float *__restrict__ W1, const float *__restrict__ b1, const float *__restrict__ W2,
                                        const float *__restrict__ b2, float *__restrict__ wp) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < NCA_HID * NCA_K1) {
        const int k = t / NCA_K1, f = t % NCA_K1;
        const float *row = W1 + (size_t)k * 3 * NCA_N;
        float v;
        if (f < 8) v = row[2 * f] + row[2 * f + 1];
        else if (f < 16) v = row[16 + 2 * (f - 8)] + row[16 + 2 * (f - 8) + 1];
        else v = row[32 + (f - 16)];
        wp[WP_W1F + t] = v;
    }
    if (t < NCA_HID) wp[WP_B1 + t] = b1[t];
    if (t < NCA_HID * NCA_N) {
        const int k = t / NCA_N, o = t % NCA_N;
        wp[WP_W2T + t] = W2[(size_t)o * NCA_HID + k];
    }
    if (t < NCA_N) wp[WP_B2 + t] = b2[t];
}

}  // namespace b200ca

using namespace b200ca;

static int nca_dims_ok(int n, int hid) {
    if (n != NCA_N || hid != NCA_HID) {
        set_error("nca: only n_states=%d, n_hidden=%d is built (got %d, %d)", NCA_N, NCA_HID, n, hid);
        return B200CA_E_UNSUPPORTED;
    }
    return 0;
}

extern "C" int64_t b200ca_nca_wprep_bytes(int n, int hid) { return (n == NCA_N && hid == NCA_HID) ? WP_TOTAL_BYTES : -1; }

extern "C" int64_t b200ca_nca_scratch_bytes(int B, int H, int W) {
    if (B < 1 || H < 1 || W < 1) return -1;
    return (int64_t)2 * B * H * ceil_div(W, 32) * sizeof(uint32_t);
}

extern "C" int b200ca_nca_prepare_weights(const float *W1, const float *b1, const float *W2, const float *b2, void *Wprep, int n,
                                          int hid, void *stream) {
    int rc = nca_dims_ok(n, hid);
    if (rc) return rc;
    B200CA_REQUIRE(W1 && b1 && W2 && b2 && Wprep, "nca_prepare_weights: null pointer");
    cudaStream_t st = as_stream(stream);
    B200CA_CUDA(cudaMemsetAsync(Wprep, 0, WP_TOTAL_BYTES, st));
    nca_prepare_simt_kernel<<<ceil_div(NCA_HID * NCA_K1, 256), 256, 0, st>>>(W1, b1, W2, b2, reinterpret_cast<float *>(Wprep));
    B200CA_LAUNCH_CHECK("nca_prepare_simt_kernel");
    return nca_prepare_tc(W1, b1, W2, b2, Wprep, st);
}

extern "C" int b200ca_nca_step(const float *state_in, float *state_out, const void *Wprep, const uint8_t *update_mask,
                               uint64_t seed, uint64_t step_index, void *scratch, int B, int n, int hid, int H, int W, int impl,
                               void *stream) {
    return b200ca_nca_step_shard(state_in, state_out, Wprep, update_mask, seed, step_index, 0, scratch, B, n, hid, H, W, impl, stream);
}

extern "C" int b200ca_nca_step_shard(const float *state_in, float *state_out, const void *Wprep, const uint8_t *update_mask,
                                     uint64_t seed, uint64_t step_index, uint64_t first_world, void *scratch, int B, int n, int hid,
                                     int H, int W, int impl, void *stream) {
    int rc = nca_dims_ok(n, hid);
    if (rc) return rc;
    B200CA_REQUIRE(state_in && state_out && Wprep && scratch, "nca_step: null pointer");
    B200CA_REQUIRE(state_in != state_out, "nca_step: in-place update is not supported");
    B200CA_REQUIRE(B >= 1 && B <= 65535 && H >= 1 && W >= 1 && H <= 65535 * 8, "nca_step: bad shape B=%d H=%d W=%d", B, H, W);
    cudaStream_t st = as_stream(stream);
    NcaArgs a;
    a.in = state_in;
    a.out = state_out;
    a.wp = reinterpret_cast<const float *>(Wprep);
    a.mask = update_mask;
    a.B = B;
    a.H = H;
    a.W = W;
    a.wpr = ceil_div(W, 32);
    a.alive_bits = reinterpret_cast<uint32_t *>(scratch);
    a.nz_bits = a.alive_bits + (size_t)B * H * a.wpr;
    a.seed = seed;
    a.step = step_index;
    a.cell0 = first_world * (uint64_t)H * (uint64_t)W;
    if (impl == B200CA_NCA_SIMT) {
        dim3 grid(a.wpr, ceil_div(H, (SIMT_THREADS / 32) * SIMT_ROWS), B);
        nca_update_simt_kernel<<<grid, SIMT_THREADS, 0, st>>>(a);
        B200CA_LAUNCH_CHECK("nca_update_simt_kernel");
        note_variant("nca_update_simt");
    } else if (impl == B200CA_NCA_TC) {
        rc = nca_step_tc(a, st);
        if (rc) return rc;
    } else {
        B200CA_REQUIRE(false, "nca_step: unknown impl %d", impl);
    }
    dim3 g2(ceil_div(W, 128), H, B);
    B200CA_REQUIRE(H <= 65535, "nca_step: H too large for the life pass");
    B200CA_CUDA(launch_pdl(nca_life_kernel, g2, dim3(128), (size_t)0, st, state_out, (const uint32_t *)a.alive_bits, (const uint32_t *)a.nz_bits, B, H, W, a.wpr));
    B200CA_LAUNCH_CHECK("nca_life_kernel");
    return 0;
}

extern "C" int b200ca_nca_run_host(const float *h_state_in, float *h_state_out, const void *Wprep, const uint8_t *h_update_mask,
                                   uint64_t seed, int B, int n, int hid, int H, int W, int impl, int nsteps) {
    int rc = nca_dims_ok(n, hid);
    if (rc) return rc;
    B200CA_REQUIRE(h_state_in && h_state_out && Wprep && nsteps >= 0, "nca_run_host: bad arguments");
    const size_t sbytes = (size_t)B * n * H * W * sizeof(float), mbytes = (size_t)B * H * W;
    float *d_a = nullptr, *d_b = nullptr;
    uint8_t *d_m = nullptr;
    void *d_s = nullptr;
    cudaStream_t st = 0;
    do {
        if ((rc = scratch_get(0, sbytes, (void **)&d_a))) break;
        if ((rc = scratch_get(1, sbytes, (void **)&d_b))) break;
        if ((rc = scratch_get(2, (size_t)b200ca_nca_scratch_bytes(B, H, W), &d_s))) break;
        if (h_update_mask) {
            if ((rc = scratch_get(3, mbytes * (nsteps > 0 ? nsteps : 1), (void **)&d_m))) break;
            if ((rc = check_cuda(cudaMemcpyAsync(d_m, h_update_mask, mbytes * nsteps, cudaMemcpyHostToDevice, st), "H2D mask"))) break;
        }
        if ((rc = check_cuda(cudaMemcpyAsync(d_a, h_state_in, sbytes, cudaMemcpyHostToDevice, st), "H2D"))) break;
        float *src = d_a, *dst = d_b;
        for (int i = 0; i < nsteps && !rc; ++i) {
            rc = b200ca_nca_step(src, dst, Wprep, d_m ? d_m + (size_t)i * mbytes : nullptr, seed, (uint64_t)i, d_s, B, n, hid, H, W,
                                 impl, st);
            float *t = src;
            src = dst;
            dst = t;
        }
        if (rc) break;
        if ((rc = check_cuda(cudaMemcpyAsync(h_state_out, src, sbytes, cudaMemcpyDeviceToHost, st), "D2H"))) break;
        rc = check_cuda(cudaStreamSynchronize(st), "sync");
    } while (0);
    return rc;
}
