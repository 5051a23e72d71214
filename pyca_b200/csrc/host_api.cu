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

// ---- pipelined host entry point --------------------------------------------------------------------------------
// A stream of HOST-resident worlds through the Lenia-family step.  Independent worlds only depend on their own state,
// so while world i steps on the SMs, world i+1 rides up the PCIe link and world i-1 rides down (separate copy engines
// per direction).  A pipe owns `depth` lanes -- stream, dense staging buffer, two framed buffers, affinity plane, FFT
// workspace, completion event; submit() deals worlds to lanes round-robin and returns as soon as the work is queued.
namespace {

struct Lane {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr, up_done = nullptr, k_done = nullptr;
    float *dense = nullptr, *a = nullptr, *b = nullptr, *aff = nullptr, *ws = nullptr;
    bool busy = false;
};

struct Pipe {
    int dev = 0, B = 0, C = 0, km = 0, H = 0, W = 0, family = 0, depth = 0, next = 0;
    float dt = 0.f, beta = 0.f, alpha = 0.f;
    bool fft = false;   // the two-kernel hybrid (needs a workspace per lane)
    int algo = 0;       // B200CA_ALGO_*
    size_t dbytes = 0;
    float *kprep = nullptr, *kfft = nullptr, *mu = nullptr, *sigma = nullptr, *weights = nullptr;
    Lane *lanes = nullptr;
    // every host-to-device copy goes through ONE stream and every device-to-host copy through another: a lane's own stream
    // may share its copy engine with another lane's opposite direction, and then the two directions queue behind each other
    cudaStream_t st_up = nullptr, st_dn = nullptr;
    int copy_mode = 3;   // bit 0: dedicated upload stream, bit 1: dedicated download stream (pageable / unmapped outputs)
};

void pipe_free(Pipe *p) {
    if (!p) return;
    if (p->lanes) {
        for (int i = 0; i < p->depth; ++i) {
            Lane &l = p->lanes[i];
            if (l.st) cudaStreamSynchronize(l.st);
            cudaFree(l.dense);
            cudaFree(l.a);
            cudaFree(l.b);
            cudaFree(l.aff);
            cudaFree(l.ws);
            if (l.done) cudaEventDestroy(l.done);
            if (l.up_done) cudaEventDestroy(l.up_done);
            if (l.k_done) cudaEventDestroy(l.k_done);
            if (l.st) cudaStreamDestroy(l.st);
        }
        delete[] p->lanes;
    }
    if (p->st_up) cudaStreamDestroy(p->st_up);
    if (p->st_dn) cudaStreamDestroy(p->st_dn);
    cudaFree(p->kprep);
    cudaFree(p->kfft);
    cudaFree(p->mu);
    cudaFree(p->sigma);
    cudaFree(p->weights);
    delete p;
    (void)cudaGetLastError();
}

int pipe_build(Pipe *p, const float *K, const float *mu, const float *sigma, const float *weights, int ks) {
    const int B = p->B, C = p->C, km = p->km, H = p->H, W = p->W;
    const size_t np = (size_t)B * C * km * C * sizeof(float);
    const size_t fbytes = (size_t)b200ca_frame_elems(B, C, H, W) * sizeof(float);
    B200CA_CUDA(cudaMalloc(&p->mu, np));
    B200CA_CUDA(cudaMalloc(&p->sigma, np));
    B200CA_CUDA(cudaMalloc(&p->weights, np));
    B200CA_CUDA(cudaMemcpy(p->mu, mu, np, cudaMemcpyDefault));
    B200CA_CUDA(cudaMemcpy(p->sigma, sigma, np, cudaMemcpyDefault));
    B200CA_CUDA(cudaMemcpy(p->weights, weights, np, cudaMemcpyDefault));
    // the reference's kernel tensor `k` may live on either side; the prepare kernels want it on the device
    const size_t kbytes = (size_t)B * C * km * C * ks * ks * sizeof(float);
    float *dK = nullptr;
    B200CA_CUDA(cudaMalloc(&dK, kbytes));
    int rc = check_cuda(cudaMemcpy(dK, K, kbytes, cudaMemcpyDefault), "K upload");
    if (!rc) {
        if (p->fft) {
            rc = check_cuda(cudaMalloc(&p->kfft, (size_t)b200ca_lenia_fft_kernel_elems(B, C, km, H, W) * sizeof(float)), "cudaMalloc");
            if (!rc) rc = b200ca_lenia_fft_prepare_kernel(dK, p->kfft, B, C, km, ks, H, W, nullptr);
        } else if (p->algo == B200CA_ALGO_MARCH) {
            rc = check_cuda(cudaMalloc(&p->kfft, (size_t)b200ca_lenia_march_kernel_elems(B, C, km) * sizeof(float)), "cudaMalloc");
            if (!rc) rc = b200ca_lenia_march_prepare_kernel(dK, p->kfft, B, C, km, ks, nullptr);
        } else {
            rc = check_cuda(cudaMalloc(&p->kprep, (size_t)b200ca_lenia_kprep_elems(B, C, km) * sizeof(float)), "cudaMalloc");
            if (!rc) rc = b200ca_lenia_prepare_kernel(dK, p->kprep, B, C, km, ks, nullptr);
        }
    }
    if (!rc) rc = check_cuda(cudaDeviceSynchronize(), "kernel preparation");
    cudaFree(dK);
    if (rc) return rc;
    B200CA_CUDA(cudaStreamCreateWithFlags(&p->st_up, cudaStreamNonBlocking));
    B200CA_CUDA(cudaStreamCreateWithFlags(&p->st_dn, cudaStreamNonBlocking));
    p->lanes = new Lane[p->depth];
    for (int i = 0; i < p->depth; ++i) {
        Lane &l = p->lanes[i];
        B200CA_CUDA(cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking));
        B200CA_CUDA(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
        B200CA_CUDA(cudaEventCreateWithFlags(&l.up_done, cudaEventDisableTiming));
        B200CA_CUDA(cudaEventCreateWithFlags(&l.k_done, cudaEventDisableTiming));
        B200CA_CUDA(cudaMalloc(&l.dense, p->dbytes));
        B200CA_CUDA(cudaMalloc(&l.a, fbytes));
        B200CA_CUDA(cudaMalloc(&l.b, fbytes));
        B200CA_CUDA(cudaMemset(l.a, 0, fbytes));
        B200CA_CUDA(cudaMemset(l.b, 0, fbytes));
        if (p->family > 0) {
            B200CA_CUDA(cudaMalloc(&l.aff, fbytes));
            B200CA_CUDA(cudaMemset(l.aff, 0, fbytes));
        }
        if (p->fft) B200CA_CUDA(cudaMalloc(&l.ws, (size_t)b200ca_lenia_fft_workspace_elems(B, C, km, H, W) * sizeof(float)));
    }
    return 0;
}

}  // namespace

extern "C" int b200ca_lenia_pipe_create(void **pipe, const float *K, const float *mu, const float *sigma, const float *weights,
                                        int B, int C, int k_mult, int ks, int H, int W, float dt, int family, float beta,
                                        float alpha, int depth) {
    B200CA_REQUIRE(pipe && K && mu && sigma && weights, "lenia_pipe_create: null pointer");
    B200CA_REQUIRE(family >= 0 && family <= 2, "lenia_pipe_create: family must be 0 (Lenia), 1 (MaCE) or 2 (MaCE-XChan)");
    B200CA_REQUIRE(depth >= 1 && depth <= 16, "lenia_pipe_create: depth must be in [1,16] (got %d)", depth);
    B200CA_REQUIRE(B >= 1 && C >= 1 && k_mult >= 1 && ks >= 1 && (ks & 1) && ks <= 31 && H >= B200CA_FRAME && W >= B200CA_FRAME,
                   "lenia_pipe_create: bad shape B=%d C=%d k_mult=%d ks=%d H=%d W=%d", B, C, k_mult, ks, H, W);
    *pipe = nullptr;
    int dev = 0;
    B200CA_CUDA(cudaGetDevice(&dev));
    Pipe *p = new Pipe();
    p->dev = dev;
    p->B = B, p->C = C, p->km = k_mult, p->H = H, p->W = W, p->family = family, p->depth = depth;
    p->dt = dt, p->beta = beta, p->alpha = alpha;
    p->algo = b200ca_lenia_pick_algo_ks(B, C * k_mult, H, W, ks);
    p->fft = p->algo == B200CA_ALGO_FFT_ROWS;
    p->dbytes = (size_t)B * C * H * W * sizeof(float);
    const int rc = pipe_build(p, K, mu, sigma, weights, ks);
    if (rc) {
        pipe_free(p);
        return rc;
    }
    *pipe = p;
    return 0;
}

extern "C" int b200ca_lenia_pipe_submit(void *pipe, const float *h_state_in, float *h_state_out, int nsteps) {
    Pipe *p = static_cast<Pipe *>(pipe);
    B200CA_REQUIRE(p && h_state_in && h_state_out && nsteps >= 0, "lenia_pipe_submit: bad arguments");
    {
        int cur = -1;
        B200CA_CUDA(cudaGetDevice(&cur));
        B200CA_REQUIRE(cur == p->dev, "lenia_pipe_submit: the pipe was created on device %d, the current device is %d", p->dev, cur);
    }
    Lane &l = p->lanes[p->next];
    p->next = (p->next + 1) % p->depth;
    if (l.busy) B200CA_CUDA(cudaEventSynchronize(l.done));  // the lane's previous world has landed in its host buffer
    const int B = p->B, C = p->C, km = p->km, H = p->H, W = p->W;
    // The result goes straight into the host buffer when it is pinned (mapped into the device's address space): the unframing
    // kernel stores over PCIe (posted writes), which saves the device-to-host copy-engine operation and its stream hop
    // (1024^2: 129 -> 114 us per world).  Reading the INPUT over PCIe from a kernel is slower than the copy engine (158 us), so
    // the host-to-device direction stays a cudaMemcpyAsync.  Pageable host buffers take the copy engine both ways.
    float *out_mapped = nullptr;
    {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, h_state_out) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
            out_mapped = static_cast<float *>(at.devicePointer);
        else
            (void)cudaGetLastError();
    }
    int copy_mode = p->copy_mode;
#ifdef B200CA_DEV
    if (const char *e = getenv("B200CA_PIPE_ZEROCOPY")) out_mapped = atoi(e) ? out_mapped : nullptr;
    if (const char *e = getenv("B200CA_PIPE_COPY_MODE")) copy_mode = atoi(e);
#endif
    if (copy_mode & 1) {
        // (the lane's previous world is complete -- cudaEventSynchronize above -- so its staging buffer is free)
        B200CA_CUDA(cudaMemcpyAsync(l.dense, h_state_in, p->dbytes, cudaMemcpyHostToDevice, p->st_up));
        B200CA_CUDA(cudaEventRecord(l.up_done, p->st_up));
        B200CA_CUDA(cudaStreamWaitEvent(l.st, l.up_done, 0));
    } else {
        B200CA_CUDA(cudaMemcpyAsync(l.dense, h_state_in, p->dbytes, cudaMemcpyHostToDevice, l.st));
    }
    int rc = b200ca_frame_load(l.dense, l.a, B, C, H, W, 1, l.st);
    float *src = l.a, *dst = l.b;
    for (int i = 0; i < nsteps && !rc; ++i) {
        float *conv_out = p->family == 0 ? dst : l.aff;
        const int mode = p->family == 0 ? B200CA_MODE_LENIA : B200CA_MODE_AFF;
        rc = p->fft ? b200ca_lenia_step_fft(src, conv_out, p->kfft, p->mu, p->sigma, p->weights, l.ws, B, C, km, H, W, p->dt, mode, 1, l.st)
             : p->algo == B200CA_ALGO_MARCH
                 ? b200ca_lenia_step_march(src, conv_out, p->kfft, p->mu, p->sigma, p->weights, B, C, km, H, W, p->dt, mode, 1, 0, 0, l.st)
                    : b200ca_lenia_step(src, conv_out, p->kprep, p->mu, p->sigma, p->weights, B, C, km, H, W, p->dt, mode, 1, 0, 0, l.st);
        if (!rc && p->family > 0)
            rc = b200ca_mace_redistribute(src, l.aff, dst, B, C, H, W, p->beta, p->family == 2, p->alpha, 1, l.st);
        float *t = src;
        src = dst;
        dst = t;
    }
    if (rc) return rc;
    if (out_mapped) {
        if ((rc = b200ca_frame_store(src, out_mapped, B, C, H, W, l.st))) return rc;
        B200CA_CUDA(cudaEventRecord(l.done, l.st));
    } else {
        if ((rc = b200ca_frame_store(src, l.dense, B, C, H, W, l.st))) return rc;
        if (copy_mode & 2) {
            B200CA_CUDA(cudaEventRecord(l.k_done, l.st));
            B200CA_CUDA(cudaStreamWaitEvent(p->st_dn, l.k_done, 0));
            B200CA_CUDA(cudaMemcpyAsync(h_state_out, l.dense, p->dbytes, cudaMemcpyDeviceToHost, p->st_dn));
            B200CA_CUDA(cudaEventRecord(l.done, p->st_dn));
        } else {
            B200CA_CUDA(cudaMemcpyAsync(h_state_out, l.dense, p->dbytes, cudaMemcpyDeviceToHost, l.st));
            B200CA_CUDA(cudaEventRecord(l.done, l.st));
        }
    }
    l.busy = true;
    return 0;
}

extern "C" int b200ca_lenia_pipe_drain(void *pipe) {
    Pipe *p = static_cast<Pipe *>(pipe);
    B200CA_REQUIRE(p, "lenia_pipe_drain: null pipe");
    for (int i = 0; i < p->depth; ++i)
        if (p->lanes[i].busy) {
            B200CA_CUDA(cudaEventSynchronize(p->lanes[i].done));
            p->lanes[i].busy = false;
        }
    return 0;
}

extern "C" int b200ca_lenia_pipe_destroy(void *pipe) {
    pipe_free(static_cast<Pipe *>(pipe));
    return 0;
}
