// Shared helpers of the b200ca kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b200ca.h"

namespace b200ca {

void set_error(const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);
int num_sms();
// names the kernel instantiation a dispatch just launched (thread-local, bounded log read back by b200ca_debug_variants():
// the parity tests assert through it that they exercised the instantiation they were written for)
void note_variant(const char *fmt, ...);
// grow-only per-device scratch buffers for the *_host entry points (slot 0..7); freed by b200ca_shutdown()
int scratch_get(int slot, size_t bytes, void **out);
void scratch_free_all();

#define B200CA_CUDA(call)                                  \
    do {                                                   \
        int _rc = b200ca::check_cuda((call), #call);       \
        if (_rc) return _rc;                               \
    } while (0)

#define B200CA_LAUNCH_CHECK(name)                                      \
    do {                                                               \
        int _rc = b200ca::check_cuda(cudaPeekAtLastError(), name);     \
        if (_rc) return _rc;                                           \
    } while (0)

#define B200CA_REQUIRE(cond, ...)                   \
    do {                                            \
        if (!(cond)) {                              \
            b200ca::set_error(__VA_ARGS__);         \
            return B200CA_E_BADARG;                 \
        }                                           \
    } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Philox4x32-10 (Salmon et al. 2011): counter-based generator used for synthetic worlds and
// for the in-kernel NCA update mask.
struct Philox {
    static __host__ __device__ __forceinline__ void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(M0, c[0]), hi1 = __umulhi(M1, c[2]);
#else
        uint32_t hi0 = (uint32_t)(((uint64_t)M0 * c[0]) >> 32), hi1 = (uint32_t)(((uint64_t)M1 * c[2]) >> 32);
#endif
        uint32_t lo0 = M0 * c[0], lo1 = M1 * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    static __host__ __device__ __forceinline__ void gen(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t (&out)[4]) {
        uint32_t c[4] = {(uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), (uint32_t)ctr_hi, (uint32_t)(ctr_hi >> 32)};
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            round(c, k0, k1);
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};


#ifdef __CUDACC__
// Programmatic dependent launch (PDL): kernels chained on one stream.  Each kernel does its input-independent prologue,
// waits for the previous grid (griddepcontrol.wait) and immediately lets the next grid start ITS prologue
// (griddepcontrol.launch_dependents): the launch latency leaves the critical path of short kernels.
__device__ __forceinline__ void pdl_wait_then_release() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename Kern, typename... Args>
static cudaError_t launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}
#endif

}  // namespace b200ca
