// Host-buffer entry points of the Lenia family: the reference-facing call with HOST tensors
// (what bench.py's `e2e` times): H2D of the dense state, nsteps device steps, D2H of the result.
#include "common.cuh"

using namespace b200ca;

extern "C" int b200ca_lenia_run_host(const float *h_state_in, float *h_state_out, const float *Kprep, const float *mu,
                                     const float *sigma, const float *weights, int B, int C, int k_mult, int H, int W, float dt,
                                     int family, float beta, float alpha, int nsteps) {
    B200CA_REQUIRE(h_state_in && h_state_out && Kprep && mu && sigma && weights && nsteps >= 0, "lenia_run_host: bad arguments");
    B200CA_REQUIRE(family >= 0 && family <= 2, "lenia_run_host: family must be 0 (Lenia), 1 (MaCE) or 2 (MaCE-XChan)");
    const size_t dbytes = (size_t)B * C * H * W * sizeof(float);
    const size_t fbytes = (size_t)b200ca_frame_elems(B, C, H, W) * sizeof(float);
    float *d_dense = nullptr, *d_a = nullptr, *d_b = nullptr, *d_aff = nullptr;
    cudaStream_t st = 0;
    int rc = 0;
    do {
        if ((rc = scratch_get(0, dbytes, (void **)&d_dense))) break;
        if ((rc = scratch_get(1, fbytes, (void **)&d_a))) break;
        if ((rc = scratch_get(2, fbytes, (void **)&d_b))) break;
        if (family > 0 && (rc = scratch_get(3, fbytes, (void **)&d_aff))) break;
        if ((rc = check_cuda(cudaMemcpyAsync(d_dense, h_state_in, dbytes, cudaMemcpyHostToDevice, st), "H2D"))) break;
        if ((rc = b200ca_frame_load(d_dense, d_a, B, C, H, W, 1, st))) break;
        float *src = d_a, *dst = d_b;
        for (int i = 0; i < nsteps && !rc; ++i) {
            if (family == 0) {
                rc = b200ca_lenia_step(src, dst, Kprep, mu, sigma, weights, B, C, k_mult, H, W, dt, B200CA_MODE_LENIA, 1, 0, 0, st);
            } else {
                rc = b200ca_lenia_step(src, d_aff, Kprep, mu, sigma, weights, B, C, k_mult, H, W, dt, B200CA_MODE_AFF, 1, 0, 0, st);
                if (!rc) rc = b200ca_mace_redistribute(src, d_aff, dst, B, C, H, W, beta, family == 2, alpha, 1, st);
            }
            float *t = src;
            src = dst;
            dst = t;
        }
        if (rc) break;
        if ((rc = b200ca_frame_store(src, d_dense, B, C, H, W, st))) break;
        if ((rc = check_cuda(cudaMemcpyAsync(h_state_out, d_dense, dbytes, cudaMemcpyDeviceToHost, st), "D2H"))) break;
        rc = check_cuda(cudaStreamSynchronize(st), "sync");
    } while (0);
    return rc;
}
