// Developer probe: isolates TMA tile-load variants (run each variant in its own process).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../pyca_b200/csrc/tma.cuh"
using namespace b200ca;
namespace b200ca { void set_error(const char*f,...){} int check_cuda(cudaError_t e,const char*w){ if(e){printf("cuda err %s: %s\n",w,cudaGetErrorString(e));} return (int)e;} int num_sms(){return 148;} }

template<int BOXW, int BOXH, bool PREFETCH>
__global__ void probe(const CUtensorMap* tmap, float* out, int x, int y) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + ((BOXW*BOXH*4 + 127)/128*128));
    if (threadIdx.x == 0) {
        if (PREFETCH) tma_prefetch_desc(tmap);
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, BOXW*BOXH*4);
        tma_load_3d(tile, tmap, bar, x, y, 0);
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < BOXW*BOXH; i += blockDim.x) out[i] = tile[i];
}

template<int BOXW, int BOXH, bool PREFETCH>
int run(int pitch, int rows, int X = 1, int Y = 1) {
    std::vector<float> h((size_t)pitch*rows);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o; cudaMalloc(&d, h.size()*4); cudaMalloc(&o, BOXW*BOXH*4);
    cudaMemcpy(d, h.data(), h.size()*4, cudaMemcpyHostToDevice);
    const CUtensorMap* dtm = nullptr;
    int rc = get_tensor_map_3d(&dtm, d, pitch, rows, 1, pitch, (uint64_t)pitch*rows, BOXW, BOXH);
    printf("encode rc=%d\n", rc);
    if (rc) return rc;
    size_t smem = (BOXW*BOXH*4 + 127)/128*128 + 64;
    cudaFuncSetAttribute(probe<BOXW,BOXH,PREFETCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<BOXW,BOXH,PREFETCH><<<1, 64, smem>>>(dtm, o, X, Y);
    cudaError_t e = cudaDeviceSynchronize();
    printf("box %dx%d prefetch %d at (%d,%d): %s\n", BOXW, BOXH, (int)PREFETCH, X, Y, cudaGetErrorString(e));
    if (!e) {
        std::vector<float> r(BOXW*BOXH);
        cudaMemcpy(r.data(), o, r.size()*4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int yy = 0; yy < BOXH; ++yy) for (int xx = 0; xx < BOXW; ++xx) {
            float exp = (yy+Y < rows && xx+X < pitch) ? (float)((yy+Y)*pitch + xx+X) : 0.f;
            if (r[yy*BOXW+xx] != exp) bad++;
        }
        printf("  mismatches: %d\n", bad);
    }
    return (int)e;
}

__global__ void probe_mbar_only(float* out) {
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&bar)) : "memory");
    mbar_wait(&bar, 0);
    out[threadIdx.x] = 1.f;
}
__global__ void probe_bulk1d(const float* in, float* out) {
    __shared__ __align__(128) float tile[1024];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, 4096);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(tile)), "l"(in), "r"(4096), "r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = tile[i];
}
template<int BOXW, int BOXH>
__global__ void probe_gdesc(const CUtensorMap* tmap, float* out, int x, int y) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + ((BOXW*BOXH*4 + 127)/128*128));
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(bar, BOXW*BOXH*4); tma_load_3d(tile, tmap, bar, x, y, 0); }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < BOXW*BOXH; i += blockDim.x) out[i] = tile[i];
}
int run_simple(int v) {
    float *d, *o; cudaMalloc(&d, 1<<20); cudaMalloc(&o, 1<<20); cudaMemset(d, 0, 1<<20);
    if (v == 10) probe_mbar_only<<<1, 64>>>(o);
    if (v == 11) probe_bulk1d<<<1, 64>>>(d, o);
    if (v == 12) {
        const CUtensorMap* dtm = nullptr; int rc = get_tensor_map_3d(&dtm, d, 256, 256, 1, 256, 65536, 64, 32); printf("encode rc=%d\n", rc);
        probe_gdesc<64,32><<<1, 64, 64*32*4 + 64>>>(dtm, o, 0, 0);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("simple variant %d: %s\n", v, cudaGetErrorString(e));
    return (int)e;
}
int main(int argc, char** argv) {
    if (argc > 1 && atoi(argv[1]) >= 10 && atoi(argv[1]) < 20) return run_simple(atoi(argv[1]));
    int v = argc > 1 ? atoi(argv[1]) : 0;
    switch (v) {
        case 0: return run<100, 46, true>(1056, 1056);
        case 20: return run<100, 46, false>(1056, 1056, atoi(argv[2]), atoi(argv[3]));
        case 21: return run<100, 46, true>(1056, 1056, atoi(argv[2]), atoi(argv[3]));
        case 1: return run<100, 46, false>(1056, 1056);
        case 2: return run<96, 46, false>(1056, 1056);
        case 3: return run<128, 46, false>(1056, 1056);
        case 4: return run<64, 32, false>(1056, 1056);
        case 5: return run<100, 94, true>(96, 80);
    }
    return 0;
}
