// Library-wide plumbing: error text, device queries, framed-layout helpers.
#include <stdarg.h>

#include <algorithm>
#include "common.cuh"

namespace b200ca {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static thread_local char g_variants[1024] = "";

void note_variant(const char *fmt, ...) {
    char one[128];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(one, sizeof(one), fmt, ap);
    va_end(ap);
    const size_t have = strlen(g_variants), add = strlen(one);
    if (have + add + 2 >= sizeof(g_variants)) return;  // bounded: the log is only read by tests, right after a reset
    if (have) g_variants[have] = ';';
    memcpy(g_variants + have + (have ? 1 : 0), one, add + 1);
}

int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    (void)cudaGetLastError();  // clear the sticky-free error so the next call reports its own
    return (int)e;
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

struct ScratchSlot {
    void *p;
    size_t bytes;
};
static ScratchSlot g_scratch[64][8] = {};

int scratch_get(int slot, size_t bytes, void **out) {
    int dev = 0;
    B200CA_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || slot < 0 || slot >= 8) {
        set_error("scratch_get: bad device/slot");
        return B200CA_E_BADARG;
    }
    ScratchSlot &s = g_scratch[dev][slot];
    if (s.bytes < bytes) {
        if (s.p) {
            B200CA_CUDA(cudaDeviceSynchronize());
            cudaFree(s.p);
            s.p = nullptr;
            s.bytes = 0;
        }
        B200CA_CUDA(cudaMalloc(&s.p, bytes));
        s.bytes = bytes;
    }
    *out = s.p;
    return 0;
}

void scratch_free_all() {
    int cur = 0;
    cudaGetDevice(&cur);
    for (int d = 0; d < 64; ++d)
        for (int k = 0; k < 8; ++k)
            if (g_scratch[d][k].p) {
                cudaSetDevice(d);
                cudaDeviceSynchronize();
                cudaFree(g_scratch[d][k].p);
                g_scratch[d][k].p = nullptr;
                g_scratch[d][k].bytes = 0;
            }
    cudaSetDevice(cur);
    (void)cudaGetLastError();
}

// ---- framed layout ---------------------------------------------------------------------------
// frame cell (fy, fx) in [0, H+2F) x [0, W+2F) holds interior cell ((fy-F) mod H, (fx-F) mod W)
__global__ void frame_load_kernel(const float *__restrict__ dense, float *__restrict__ framed, int planes, int H, int W,
                                  int pitch, int wrap_rows) {
    const int F = B200CA_FRAME;
    const int fx = blockIdx.x * blockDim.x + threadIdx.x;
    const int fy = blockIdx.y;
    const int p = blockIdx.z;
    if (fx >= W + 2 * F) return;
    int y = fy - F, x = fx - F;
    if (!wrap_rows && (y < 0 || y >= H)) return;
    y = (y % H + H) % H;
    x = (x % W + W) % W;
    framed[((size_t)p * (H + 2 * F) + fy) * pitch + fx] = dense[((size_t)p * H + y) * W + x];
}

__global__ void frame_store_kernel(const float *__restrict__ framed, float *__restrict__ dense, int planes, int H, int W, int pitch) {
    const int F = B200CA_FRAME;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int p = blockIdx.z;
    if (x >= W) return;
    dense[((size_t)p * H + y) * W + x] = framed[((size_t)p * (H + 2 * F) + y + F) * pitch + x + F];
}

// the same, four cells per thread (W % 4 == 0, 16-byte aligned dense rows): 512 contiguous bytes per warp -- the form the host
// pipe uses to store a result straight into a pinned host buffer (posted PCIe writes, no copy-engine hop)
__global__ void frame_store_vec4_kernel(const float *__restrict__ framed, float *__restrict__ dense, int planes, int H, int W, int pitch) {
    const int F = B200CA_FRAME;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    const int p = blockIdx.z;
    if (x >= W) return;
    const float4 v = *reinterpret_cast<const float4 *>(framed + ((size_t)p * (H + 2 * F) + y + F) * pitch + x + F);
    *reinterpret_cast<float4 *>(dense + ((size_t)p * H + y) * W + x) = v;
}

// copies interior edges into the frame (everything outside the interior)
__global__ void frame_refresh_kernel(float *__restrict__ framed, int planes, int H, int W, int pitch, int wrap_rows) {
    const int F = B200CA_FRAME;
    const int fx = blockIdx.x * blockDim.x + threadIdx.x;
    const int fy = blockIdx.y;
    const int p = blockIdx.z;
    if (fx >= W + 2 * F) return;
    int y = fy - F, x = fx - F;
    const bool in_y = (y >= 0 && y < H), in_x = (x >= 0 && x < W);
    if (in_y && in_x) return;
    if (!wrap_rows && !in_y) return;
    y = (y % H + H) % H;
    x = (x % W + W) % W;
    float *plane = framed + (size_t)p * (H + 2 * F) * pitch;
    plane[(size_t)fy * pitch + fx] = plane[(size_t)(y + F) * pitch + x + F];
}

__global__ void frame_mass_kernel(const float *__restrict__ framed, double *__restrict__ mass, int H, int W, int pitch) {
    const int F = B200CA_FRAME;
    const int p = blockIdx.y;
    const float *plane = framed + (size_t)p * (H + 2 * F) * pitch;
    double local = 0.0;
    const size_t total = (size_t)H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t y = i / W, x = i % W;
        local += (double)plane[(y + F) * pitch + x + F];
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ double sm[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sm[wid] = local;
    __syncthreads();
    if (wid == 0) {
        local = lane < (blockDim.x >> 5) ? sm[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if (lane == 0) atomicAdd(mass + p, local);
    }
}

}  // namespace b200ca

using namespace b200ca;

extern "C" int b200ca_version(void) { return B200CA_VERSION; }
extern "C" const char *b200ca_last_error(void) { return g_err; }
extern "C" const char *b200ca_debug_variants(int reset) {
    static thread_local char out[sizeof(g_variants)];
    memcpy(out, g_variants, sizeof(out));
    if (reset) g_variants[0] = 0;
    return out;
}

extern "C" int b200ca_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return B200CA_E_NODEVICE;
    cudaDeviceProp prop;
    rc = check_cuda(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties");
    if (rc) return B200CA_E_NODEVICE;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return 0;
}

extern "C" int b200ca_frame_pitch(int W) { return (W + 2 * B200CA_FRAME + 3) / 4 * 4; }
extern "C" int64_t b200ca_frame_elems(int B, int C, int H, int W) {
    return (int64_t)B * C * (H + 2 * B200CA_FRAME) * b200ca_frame_pitch(W);
}

static int frame_args_ok(const void *a, const void *b, int B, int C, int H, int W) {
    B200CA_REQUIRE(a && b, "frame: null pointer");
    B200CA_REQUIRE(B >= 1 && C >= 1 && H >= B200CA_FRAME && W >= B200CA_FRAME, "frame: need B,C>=1 and H,W>=%d (got %d,%d,%d,%d)",
                   B200CA_FRAME, B, C, H, W);
    B200CA_REQUIRE((long long)B * C <= 65535, "frame: B*C too large");
    return 0;
}

extern "C" int b200ca_frame_load(const float *dense, float *framed, int B, int C, int H, int W, int wrap_rows, void *stream) {
    int rc = frame_args_ok(dense, framed, B, C, H, W);
    if (rc) return rc;
    const int F = B200CA_FRAME;
    dim3 grid(ceil_div(W + 2 * F, 256), H + 2 * F, B * C);
    frame_load_kernel<<<grid, 256, 0, as_stream(stream)>>>(dense, framed, B * C, H, W, b200ca_frame_pitch(W), wrap_rows);
    B200CA_LAUNCH_CHECK("frame_load_kernel");
    return 0;
}

extern "C" int b200ca_frame_store(const float *framed, float *dense, int B, int C, int H, int W, void *stream) {
    int rc = frame_args_ok(dense, framed, B, C, H, W);
    if (rc) return rc;
    bool vec = W % 4 == 0 && (((uintptr_t)dense | (uintptr_t)framed) & 15) == 0;
#ifdef B200CA_DEV
    if (const char *e = getenv("B200CA_STORE_VEC")) vec = vec && atoi(e) != 0;
#endif
    if (vec) {
        dim3 grid(ceil_div(W / 4, 256), H, B * C);
        frame_store_vec4_kernel<<<grid, 256, 0, as_stream(stream)>>>(framed, dense, B * C, H, W, b200ca_frame_pitch(W));
        B200CA_LAUNCH_CHECK("frame_store_vec4_kernel");
        return 0;
    }
    dim3 grid(ceil_div(W, 256), H, B * C);
    frame_store_kernel<<<grid, 256, 0, as_stream(stream)>>>(framed, dense, B * C, H, W, b200ca_frame_pitch(W));
    B200CA_LAUNCH_CHECK("frame_store_kernel");
    return 0;
}

extern "C" int b200ca_frame_refresh(float *framed, int B, int C, int H, int W, int wrap_rows, void *stream) {
    int rc = frame_args_ok(framed, framed, B, C, H, W);
    if (rc) return rc;
    const int F = B200CA_FRAME;
    dim3 grid(ceil_div(W + 2 * F, 256), H + 2 * F, B * C);
    frame_refresh_kernel<<<grid, 256, 0, as_stream(stream)>>>(framed, B * C, H, W, b200ca_frame_pitch(W), wrap_rows);
    B200CA_LAUNCH_CHECK("frame_refresh_kernel");
    return 0;
}

extern "C" int b200ca_frame_mass(const float *framed, double *mass, int B, int C, int H, int W, void *stream) {
    int rc = frame_args_ok(framed, mass, B, C, H, W);
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    B200CA_CUDA(cudaMemsetAsync(mass, 0, sizeof(double) * B * C, st));
    int bx = (int)std::min<long long>(ceil_div64((long long)H * W, 256 * 8), 256);
    dim3 grid(bx, B * C);
    frame_mass_kernel<<<grid, 256, 0, st>>>(framed, mass, H, W, b200ca_frame_pitch(W));
    B200CA_LAUNCH_CHECK("frame_mass_kernel");
    return 0;
}
