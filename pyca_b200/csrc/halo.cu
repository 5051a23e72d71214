// Peer-to-peer halo exchange of a row slab as ONE kernel per rank (sm_100a, NVLink / NVSwitch peer mappings).
//
// Used by the Game-of-Life slabs (pyca_b200/slab.py: GolSlabP2P), which exchange `g` ghost rows every `g` generations: an
// NCCL send/recv pair per neighbour costs ~90 us of host and device time per exchange against ~1.5 ms of generations; this
// kernel stores the boundary rows straight into the neighbours' ghost rows and synchronises the two ranks through four
// counters in (symmetric) device memory:
//   A. "my generations are done"      -> bump the neighbours' `ready` counters (they may now overwrite my ghost rows),
//   B. wait for my own `ready` counters (the neighbours finished reading THEIR ghost rows),
//   C. copy my first / last rows into the up / down neighbour's bottom / top ghost rows (16-byte stores over NVLink),
//   D. fence, then the last block bumps the neighbours' `arrived` counters,
//   E. the last block waits for my own `arrived` counters: when the kernel ends, my ghost rows are current.
// Every rank issues the same sequence of exchanges (epoch = 1, 2, ...).  Waits are bounded (~4 s): a lost neighbour raises
// *err_flag instead of hanging the GPU.
#include <algorithm>

#include "common.cuh"

namespace b200ca {

struct HaloArgs {
    const int4 *src_up, *src_dn;   // my first / last owned rows
    int4 *dst_up, *dst_dn;         // the up neighbour's bottom ghost rows / the down neighbour's top ghost rows (peer mappings)
    long long n16;                 // 16-byte words per block of rows
    int nplanes;                   // blocks per direction (the planes of a framed Lenia slab), `stride16` 16-byte words apart
    long long stride16;
    unsigned *up_ready, *dn_ready;       // neighbours' counters bumped in A
    const unsigned *my_ready;            // [0] bumped by the up neighbour, [1] by the down neighbour
    unsigned *up_arrived, *dn_arrived;   // neighbours' counters bumped in D
    const unsigned *my_arrived;          // [0] top ghost rows arrived, [1] bottom ghost rows arrived
    unsigned epoch;
    unsigned *ticket;              // local scratch counter (zero between calls)
    unsigned *err_flag;
};

__device__ __forceinline__ void sys_add(unsigned *p) { asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ bool sys_wait(const unsigned *p, unsigned target, unsigned *err) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        if ((int)(v - target) >= 0) return true;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ull) {
            if (err) atomicExch(err, 1u);
            return false;
        }
        __nanosleep(100);
    }
}

__global__ void __launch_bounds__(256) halo_push_kernel(const HaloArgs a) {
    __shared__ unsigned s_last;
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // A
        sys_add(a.up_ready);
        sys_add(a.dn_ready);
    }
    if (threadIdx.x == 0) {  // B
        sys_wait(a.my_ready, a.epoch, a.err_flag);
        sys_wait(a.my_ready + 1, a.epoch, a.err_flag);
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;  // C
    for (int p = 0; p < a.nplanes; ++p) {
        const long long o = p * a.stride16;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n16; i += stride) {
            a.dst_up[o + i] = a.src_up[o + i];
            a.dst_dn[o + i] = a.src_dn[o + i];
        }
    }
    __threadfence_system();  // D
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        *a.ticket = 0;
        sys_add(a.up_arrived);
        sys_add(a.dn_arrived);
        sys_wait(a.my_arrived, a.epoch, a.err_flag);  // E
        sys_wait(a.my_arrived + 1, a.epoch, a.err_flag);
    }
}

}  // namespace b200ca

using namespace b200ca;

extern "C" int b200ca_halo_push_planes(const void *src_up, void *dst_up, const void *src_dn, void *dst_dn, int64_t nbytes, int nplanes,
                                       int64_t plane_stride_bytes, uint32_t *up_ready, uint32_t *dn_ready, const uint32_t *my_ready,
                                       uint32_t *up_arrived, uint32_t *dn_arrived, const uint32_t *my_arrived, uint32_t epoch,
                                       uint32_t *ticket, uint32_t *err_flag, void *stream) {
    B200CA_REQUIRE(src_up && dst_up && src_dn && dst_dn && up_ready && dn_ready && my_ready && up_arrived && dn_arrived && my_arrived && ticket,
                   "halo_push: null pointer");
    B200CA_REQUIRE(nbytes > 0 && nbytes % 16 == 0, "halo_push: the row blocks must be multiples of 16 bytes (got %lld)", (long long)nbytes);
    B200CA_REQUIRE(nplanes >= 1 && (nplanes == 1 || (plane_stride_bytes >= nbytes && plane_stride_bytes % 16 == 0)),
                   "halo_push: bad plane count / stride (%d, %lld)", nplanes, (long long)plane_stride_bytes);
    B200CA_REQUIRE((((uintptr_t)src_up | (uintptr_t)dst_up | (uintptr_t)src_dn | (uintptr_t)dst_dn) & 15) == 0, "halo_push: 16-byte alignment");
    HaloArgs a;
    a.src_up = static_cast<const int4 *>(src_up);
    a.src_dn = static_cast<const int4 *>(src_dn);
    a.dst_up = static_cast<int4 *>(dst_up);
    a.dst_dn = static_cast<int4 *>(dst_dn);
    a.n16 = nbytes / 16;
    a.nplanes = nplanes;
    a.stride16 = plane_stride_bytes / 16;
    a.up_ready = up_ready;
    a.dn_ready = dn_ready;
    a.my_ready = my_ready;
    a.up_arrived = up_arrived;
    a.dn_arrived = dn_arrived;
    a.my_arrived = my_arrived;
    a.epoch = epoch;
    a.ticket = ticket;
    a.err_flag = err_flag;
    const int blocks = (int)std::min<long long>(32, (a.n16 + 255) / 256);
    halo_push_kernel<<<blocks, 256, 0, as_stream(stream)>>>(a);
    B200CA_LAUNCH_CHECK("halo_push_kernel");
    note_variant("halo_push<blocks=%d,planes=%d>", blocks, nplanes);
    return 0;
}

extern "C" int b200ca_halo_push(const void *src_up, void *dst_up, const void *src_dn, void *dst_dn, int64_t nbytes, uint32_t *up_ready,
                                uint32_t *dn_ready, const uint32_t *my_ready, uint32_t *up_arrived, uint32_t *dn_arrived,
                                const uint32_t *my_arrived, uint32_t epoch, uint32_t *ticket, uint32_t *err_flag, void *stream) {
    return b200ca_halo_push_planes(src_up, dst_up, src_dn, dst_dn, nbytes, 1, 0, up_ready, dn_ready, my_ready, up_arrived, dn_arrived, my_arrived,
                                   epoch, ticket, err_flag, stream);
}
