// Neural-CA update pass on the 5th-generation tensor cores (tcgen05 + TMEM) for sm_100a.
//
// Same arithmetic as nca_update_simt_kernel (nca.cu) -- perception (nca.py:328-344), Linear-ReLU-Linear
// (nca.py:205-208), sigmoid * update mask, pre-update life mask (nca.py:226-237, :256-263) -- but the per-cell MLP,
// a real dense contraction, runs as two tcgen05.mma GEMMs per 128-cell tile:
//     D1[128 x 128] = P[128 x 32] * W1f^T          (folded first layer, K = 32)
//     D2[128 x 16]  = relu(D1 + b1)[128 x 128] * W2^T   (K = 128)
// with fp32 accumulators in tensor memory.  fp32-grade accuracy (the 1e-5 single-step parity bound rules out
// plain tf32/bf16 operands) comes from a 2-way tf32 split of BOTH operands and three MMAs per
// product, hi*hi + hi*lo + lo*hi (the dropped lo*lo term is 2^-22 relative).
//
// One CTA = 128 threads = 128 tile rows: thread t owns cell (warp row, lane column) of a 32x4 cell block, which is
// also TMEM lane t of both accumulators, so every stage is "one thread, one cell":
//   perception in registers -> hi/lo split -> A operand in shared memory (no-swizzle K-major core matrices:
//   [k-chunk of 4][row][4 floats], conflict-free 16-byte stores) -> thread 0 issues the MMAs, tcgen05.commit ->
//   mbarrier -> tcgen05.ld of the thread's own accumulator row -> bias/ReLU/split -> A operand of GEMM 2, fed in
//   16-column chunks so the second GEMM's MMAs trail the epilogue of the first -> sigmoid, masks, coalesced NCHW
//   stores + the two bit planes for the life-mask pass.
// A CTA holds NG = 4 independent 128-thread groups (16 warps per SM) that share ONE copy of the weight images in shared
// memory; each group walks its own tiles with its own A buffers, mbarriers, named barrier and 128 TMEM columns (D2
// reuses D1's first 16 columns, which the first epilogue chunk has drained before the first GEMM-2 MMA is issued), so
// the CUDA-core phases of three groups overlap the tensor phases of the fourth.  CTAs are persistent (grid = #SMs).
#include <stdlib.h>

#include <algorithm>

#include "nca.cuh"
#include "tma.cuh"

namespace b200ca {

// ---- Wprep tensor-core section (byte offsets from WP_TC_OFF) -------------------------------------------------
// B1hi/B1lo: [W1f | b1 | 0] as [10 k-chunks][128 n][4 floats] (K = 40: column 32 is the bias, fed by a constant-1
// feature, columns 33..39 are zero);  B2hi/B2lo: W2 as [32 k-chunks][16 n][4 floats]
constexpr int K1P = 40;
constexpr int TC_B1HI = 0;
constexpr int TC_B1LO = TC_B1HI + NCA_HID * K1P * 4;      // 20 KB each
constexpr int TC_B2HI = TC_B1LO + NCA_HID * K1P * 4;
constexpr int TC_B2LO = TC_B2HI + NCA_N * NCA_HID * 4;    // 8 KB each
constexpr int TC_BIAS1 = TC_B2LO + NCA_N * NCA_HID * 4;   // 128 floats
constexpr int TC_BIAS2 = TC_BIAS1 + NCA_HID * 4;          // 16 floats
constexpr int TC_WBYTES = TC_BIAS2 + NCA_N * 4;           // 57920
static_assert(TC_WBYTES <= WP_TC_BYTES, "Wprep tensor-core section overflows");

// ---- shared memory plan ----------------------------------------------------------------------------------------
constexpr int SM_W = 0;                                   // weight images + biases (TC_WBYTES, padded)
constexpr int SM_A1HI = 57 * 1024;                        // [10][128][4] floats = 20 KB
constexpr int SM_A1LO = SM_A1HI + 20 * 1024;
// A2 chunk buffers ([4][128][4] floats = 8 KB hi + 8 KB lo per 16 hidden columns, double-buffered = 32 KB) alias the
// start of A1: GEMM 1 has consumed A1 (bar1) before the first chunk is written, and the next tile rewrites A1 only
// after the last GEMM-2 commits (bar2)
constexpr int SM_A2 = SM_A1HI;            // buffer b at SM_A2 + b * 16 KB: hi then lo
constexpr int A2_BUF = 16 * 1024;
constexpr int NG = 4;                                     // 128-thread groups per CTA
constexpr int SM_GROUP = 40 * 1024;                       // per-group A region (A1hi + A1lo; A2 buffers alias it)
constexpr int SM_BAR = SM_A1HI + NG * SM_GROUP;           // mbarriers (6 per group) + tmem address
constexpr int SM_TOTAL = SM_BAR + 256;                    // 217 KB + 256: one CTA per SM
constexpr int NTHREADS = 128 * NG + 32 * NG;              // four worker groups + one MMA-issue warp per group
static_assert(SM_TOTAL + 1024 <= 227 * 1024, "CTA must fit in one SM");
static_assert(TC_WBYTES <= SM_A1HI, "weights overflow their smem region");

constexpr int TMEM_COLS = 512;  // 128 columns per group: D1 = [0,128), D2 = [0,16) once the first chunk is drained
constexpr int D2_COL = 0;
constexpr int H2CHUNK = 16;     // hidden columns per GEMM-2 chunk (two K=8 steps)

// tf32 split: hi = x truncated to 10 explicit mantissa bits (low 13 bits zero, so the tensor core reads it exactly);
// lo = x - hi (exact in fp32, |lo| < 2^-10 |x|; the tensor core keeps its top 11 bits: error < 2^-20 |x|, and the dropped
// lo*lo product is < 2^-20 |a||b| -- both far inside the 1e-5 parity bound, measured 5e-7).  Truncation instead of
// round-to-nearest saves the add of the rounding form: one LOP3 + one FADD per element (cvt.rna.tf32.f32 expands to ~6
// SASS instructions on sm_100); the WEIGHT images keep the rounded split (prepared once).
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE: 8x(16 B) core matrices; LBO = byte stride between the two
// 16-byte K chunks of one MMA, SBO = byte stride between 8-row groups (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) |
           (1ull << 46);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B tf32 K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// one MMA with the A / B descriptors given as (constant high word, low word): the low word carries the 16-byte-granular
// shared-memory address, so stepping through an operand is one integer add
template <bool ACC>
__device__ __forceinline__ void mma_tf32_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t b_hi, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "n"(ACC ? 1 : 0)
        : "memory");
}
// true in exactly one lane of a converged warp (elect.sync): the code under it is provably single-lane, so ptxas moves the
// MMA operands to uniform registers directly instead of wrapping every tcgen05.mma in a lane-election loop
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// sigmoid(x) = 1 / (1 + 2^(-x log2 e)) with the SFU ex2/rcp approximations (relative error < 4e-7, far inside the
// 1e-5 parity bound; the reference's torch.sigmoid is itself only ~1 ulp accurate)
__device__ __forceinline__ float sigmoid_rn(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    return r;
}

__device__ __forceinline__ void group_sync(int gid) { asm volatile("bar.sync %0, 128;" ::"r"(gid + 1) : "memory"); }

// Roles: 16 worker warps (four 128-thread groups, one tile each: perception, ReLU / splits, output) and one MMA-issue warp
// PER GROUP.  Hand-offs are mbarriers in both directions -- workers arrive on full1 / full2[b] (128 arrivals) when an A
// operand is in shared memory, the group's MMA warp waits for it, issues the tcgen05.mma burst and commits to bar1 /
// bar2[b], which the workers wait on -- so no worker warp ever issues an MMA or waits at a group-wide bar.sync: with the
// issue inside worker warp 0 (round 1) 27 % of the warp samples sat at the group barrier after every operand store and at
// the re-convergence behind the issuing lane.  (ONE issue warp polling all four groups was 1.8x slower: a tcgen05.commit
// tracks every earlier MMA of the issuing thread, so each group's hand-off waited for the other groups' GEMMs.)
__global__ void __launch_bounds__(NTHREADS, 1) nca_update_tc_kernel(const NcaArgs a, int tiles_x, int tiles_y) {
    extern __shared__ __align__(1024) unsigned char smem_all[];
    const bool mma_warp = threadIdx.x >= 128 * NG;
    const int gid = mma_warp ? (threadIdx.x - 128 * NG) >> 5 : threadIdx.x >> 7;   // group of this thread
    const int tid = threadIdx.x & 127;          // thread within the group = tile row = TMEM lane
    const int lane = tid & 31, warp = tid >> 5;
    unsigned char *smem = smem_all + gid * SM_GROUP;  // group view: the SM_A* offsets below land in this group's region
    float *sw = reinterpret_cast<float *>(smem_all + SM_W);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_all + SM_BAR);
    uint64_t *bar1 = bars + 6 * gid;   // GEMM 1 done (MMA warp -> workers)
    uint64_t *bar2 = bar1 + 1;         // bar2[b]: GEMM 2 chunk using A2 buffer b done
    uint64_t *full1 = bar1 + 3;        // A1 written by all 128 workers (workers -> MMA warp)
    uint64_t *full2 = bar1 + 4;        // full2[b]: A2 buffer b written
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem_all + SM_BAR + 6 * NG * 8);

    // ---- one-time setup: weights -> smem, barriers, TMEM
    {
        const float4 *src = reinterpret_cast<const float4 *>(reinterpret_cast<const unsigned char *>(a.wp) + WP_TC_OFF);
        float4 *dst = reinterpret_cast<float4 *>(sw);
        for (int t = threadIdx.x; t < TC_WBYTES / 16; t += NTHREADS) dst[t] = __ldg(src + t);
    }
    if (tid == 0 && !mma_warp) {
        mbar_init(bar1, 1);
        mbar_init(bar2, 1);
        mbar_init(bar2 + 1, 1);
        mbar_init(full1, 128);
        mbar_init(full2, 128);
        mbar_init(full2 + 1, 128);
        mbar_fence_init();
    }
    if (threadIdx.x < 32) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();  // weight images are read by the tensor core (async proxy)
    tc_fence_before();
    // programmatic dependent launch: everything above (57 KB of weights into shared memory, barriers, TMEM) overlaps the tail of the
    // previous step's life pass; the state and the bit planes are only touched below
    pdl_wait_then_release();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot + gid * 128;                   // this group's 128 columns
    const uint32_t my_tmem = tmem + ((uint32_t)(warp * 32) << 16);  // this warp's 32 TMEM lanes

    const uint32_t sA1hi = smem_u32(smem + SM_A1HI), sA1lo = smem_u32(smem + SM_A1LO);
    const uint32_t sA2 = smem_u32(smem + SM_A2);
    const uint32_t sB1hi = smem_u32(smem_all + SM_W + TC_B1HI), sB1lo = smem_u32(smem_all + SM_W + TC_B1LO);
    const uint32_t sB2hi = smem_u32(smem_all + SM_W + TC_B2HI), sB2lo = smem_u32(smem_all + SM_W + TC_B2LO);
    const float *bias2 = sw + TC_BIAS2 / 4;
    // descriptor strides: A operands [chunk][128 rows][16 B], B1 [chunk][128 n][16 B], B2 [chunk][16 n][16 B]
    constexpr uint32_t lboA = 128 * 16, lboB1 = 128 * 16, lboB2 = 16 * 16, sboA = 128, sboB1 = 128, sboB2 = 128;
    constexpr uint32_t IDESC1 = make_idesc(128), IDESC2 = make_idesc(16);
    // descriptor words: low = (address >> 4) | (LBO >> 4) << 16, high = (SBO >> 4) | version bit (see make_desc)
    constexpr uint32_t DHI_A = (sboA >> 4) | (1u << 14), DHI_B1 = (sboB1 >> 4) | (1u << 14), DHI_B2 = (sboB2 >> 4) | (1u << 14);
    const uint32_t dA1hi = ((sA1hi >> 4) & 0x3fffu) | ((lboA >> 4) << 16), dA1lo = ((sA1lo >> 4) & 0x3fffu) | ((lboA >> 4) << 16);
    const uint32_t dB1hi = ((sB1hi >> 4) & 0x3fffu) | ((lboB1 >> 4) << 16), dB1lo = ((sB1lo >> 4) & 0x3fffu) | ((lboB1 >> 4) << 16);
    const uint32_t dA2 = ((sA2 >> 4) & 0x3fffu) | ((lboA >> 4) << 16);
    const uint32_t dB2hi = ((sB2hi >> 4) & 0x3fffu) | ((lboB2 >> 4) << 16), dB2lo = ((sB2lo >> 4) & 0x3fffu) | ((lboB2 >> 4) << 16);

    const size_t plane = (size_t)a.H * a.W;
    const int ntiles = tiles_x * tiles_y * a.B;
    uint32_t ph1 = 0, ph2[2] = {0, 0};

    if (mma_warp) {
        // ---- MMA-issue warp of group `gid`: per tile the requests are [GEMM 1, chunk 0, ..., chunk 7]
        uint32_t pf1 = 0, pf2[2] = {0, 0};
        const uint32_t tm = tmem;
        for (int tile = blockIdx.x * NG + gid; tile < ntiles; tile += gridDim.x * NG) {
            mbar_wait(full1, pf1);
            pf1 ^= 1;
            if (elect_one()) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < K1P / 8; ++ks) {
                    constexpr uint32_t step = (2 * 128 * 16) >> 4;   // two 16-byte K chunks of 128 rows per k-step
                    const uint32_t ah = dA1hi + ks * step, al = dA1lo + ks * step, bh = dB1hi + ks * step, bl = dB1lo + ks * step;
                    if (ks == 0) mma_tf32_lo<false>(tm, ah, bh, DHI_A, DHI_B1, IDESC1);
                    else mma_tf32_lo<true>(tm, ah, bh, DHI_A, DHI_B1, IDESC1);
                    mma_tf32_lo<true>(tm, ah, bl, DHI_A, DHI_B1, IDESC1);
                    mma_tf32_lo<true>(tm, al, bh, DHI_A, DHI_B1, IDESC1);
                }
                mma_commit(bar1);
            }
            __syncwarp();
#pragma unroll 2
            for (int c = 0; c < NCA_HID / H2CHUNK; ++c) {
                const int bsel = c & 1;
                mbar_wait(full2 + bsel, pf2[bsel]);
                pf2[bsel] ^= 1;
                if (elect_one()) {
                    tc_fence_after();
                    const uint32_t a2hi = dA2 + ((bsel * A2_BUF) >> 4), a2lo = a2hi + ((A2_BUF / 2) >> 4);
#pragma unroll
                    for (int s2 = 0; s2 < H2CHUNK / 8; ++s2) {
                        const uint32_t offA = (s2 * 2 * 128 * 16) >> 4;
                        const uint32_t offB = ((c * (H2CHUNK / 4) + s2 * 2) * 16 * 16) >> 4;
                        const uint32_t ah = a2hi + offA, al = a2lo + offA, bh = dB2hi + offB, bl = dB2lo + offB;
                        if (s2 == 0 && c == 0) mma_tf32_lo<false>(tm + D2_COL, ah, bh, DHI_A, DHI_B2, IDESC2);
                        else mma_tf32_lo<true>(tm + D2_COL, ah, bh, DHI_A, DHI_B2, IDESC2);
                        mma_tf32_lo<true>(tm + D2_COL, ah, bl, DHI_A, DHI_B2, IDESC2);
                        mma_tf32_lo<true>(tm + D2_COL, al, bh, DHI_A, DHI_B2, IDESC2);
                    }
                    mma_commit(bar2 + bsel);
                }
                __syncwarp();
            }
        }
    } else
    for (int tile = blockIdx.x * NG + gid; tile < ntiles; tile += gridDim.x * NG) {
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
        const int x = tx * 32 + lane, y = ty * 4 + warp;
        const bool valid = x < a.W && y < a.H;
        const int xc = min(x, a.W - 1), yc = min(y, a.H - 1);
        const float *sb = a.in + (size_t)b * NCA_N * plane;

        // The next tile of this group is known: pull its 6 rows x 16 channels (one 128-byte line each when the tile row is
        // line-aligned) into L2 now, so that its perception loads find them there instead of in HBM one tile later
        {
            const int nt = tile + gridDim.x * NG;
            if (nt < ntiles && tid < 6 * NCA_N) {
                const int ntx = nt % tiles_x, nty = (nt / tiles_x) % tiles_y, nb = nt / (tiles_x * tiles_y);
                const int prow = tid % 6, pch = tid / 6;
                int py = nty * 4 - 1 + prow;
                py = py < 0 ? py + a.H : (py >= a.H ? py - a.H : py);
                const float *pa = a.in + ((size_t)nb * NCA_N + pch) * plane + (size_t)py * a.W + min(ntx * 32, a.W - 1);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
            }
        }

        // ---- perception straight from global memory (L1-cached; ~114 independent loads per thread keep the memory
        //      system busy, which a staged shared-memory tile + barrier did not with only 8 warps per SM)
        bool alive0 = false;
        {
            const int xm = xc == 0 ? a.W - 1 : xc - 1, xp = xc == a.W - 1 ? 0 : xc + 1;
            const int ym = yc == 0 ? a.H - 1 : yc - 1, yp = yc == a.H - 1 ? 0 : yc + 1;
            // Addressing: the nine neighbour offsets are per-thread 32-bit BYTE offsets inside a plane, fixed for the tile; the
            // plane base is warp-uniform and steps by `plane` per channel.  A load is then [uniform base + 32-bit lane offset]
            // -- one instruction, where per-channel 64-bit pointer arithmetic cost ~3 integer instructions per load.
            const uint32_t b00 = 4u * (ym * a.W + xm), b01 = 4u * (ym * a.W + xc), b02 = 4u * (ym * a.W + xp);
            const uint32_t b10 = 4u * (yc * a.W + xm), b11 = 4u * (yc * a.W + xc), b12 = 4u * (yc * a.W + xp);
            const uint32_t b20 = 4u * (yp * a.W + xm), b21 = 4u * (yp * a.W + xc), b22 = 4u * (yp * a.W + xp);
            auto ld = [](const char *base, uint32_t off) { return __ldg(reinterpret_cast<const float *>(base + off)); };
            float p[NCA_K1];
#pragma unroll
            for (int c = 0; c < NCA_N; ++c) {
                const char *pl = reinterpret_cast<const char *>(sb + (size_t)c * plane);
                const float ctr = ld(pl, b11);
                p[16 + c] = ctr;
                if (c < 8) {
                    const float l0 = ld(pl, b00), r0 = ld(pl, b02);
                    const float l1 = ld(pl, b10), r1 = ld(pl, b12);
                    const float l2 = ld(pl, b20), r2 = ld(pl, b22);
                    p[c] = ((r0 - l0) + 2.f * (r1 - l1) + (r2 - l2)) * 0.125f;
                    if (c == 3) {
                        const float t1 = ld(pl, b01), b1 = ld(pl, b21);
                        const float m = fmaxf(fmaxf(fmaxf(l0, r0), fmaxf(l1, r1)), fmaxf(fmaxf(l2, r2), fmaxf(t1, b1)));
                        alive0 = fmaxf(m, ctr) > 0.1f;
                    }
                } else {
                    const float t0 = ld(pl, b00), t1 = ld(pl, b01), t2 = ld(pl, b02);
                    const float b0 = ld(pl, b20), b1 = ld(pl, b21), b2 = ld(pl, b22);
                    p[c] = ((b0 - t0) + 2.f * (b1 - t1) + (b2 - t2)) * 0.125f;
                }
            }
            float4 *ahi = reinterpret_cast<float4 *>(smem + SM_A1HI), *alo = reinterpret_cast<float4 *>(smem + SM_A1LO);
#pragma unroll
            for (int kc = 0; kc < NCA_K1 / 4; ++kc) {
                float h[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split_tf32(p[4 * kc + e], h[e], l[e]);
                ahi[kc * 128 + tid] = make_float4(h[0], h[1], h[2], h[3]);
                alo[kc * 128 + tid] = make_float4(l[0], l[1], l[2], l[3]);
            }
            // bias feature: constant 1 in column 32 (chunk 8), zeros up to column 39
            ahi[8 * 128 + tid] = make_float4(1.f, 0.f, 0.f, 0.f);
            ahi[9 * 128 + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
            alo[8 * 128 + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
            alo[9 * 128 + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(full1);   // -> the MMA warp issues GEMM 1: D1 = [P | 1 | 0] * [W1f | b1 | 0]^T, K = 40 = 5 k-steps x (hi*hi + hi*lo + lo*hi)
        mbar_wait(bar1, ph1);
        ph1 ^= 1;
        tc_fence_after();

        // ---- epilogue 1 + GEMM 2 in chunks of 16 hidden units, A2 double-buffered so the MMA completion latency
        //      of chunk c-1 is off the critical path of chunk c
#pragma unroll 2
        for (int c = 0; c < NCA_HID / H2CHUNK; ++c) {
            const int bsel = c & 1;
            float v[16];
            tmem_ld16(my_tmem + c * H2CHUNK, v);
            float hh[16], hl[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) split_tf32(fmaxf(v[e], 0.f), hh[e], hl[e]);
            if (c >= 2) {  // the MMAs of chunk c-2 must have consumed this buffer
                mbar_wait(bar2 + bsel, ph2[bsel]);
                ph2[bsel] ^= 1;
            }
            float4 *ahi = reinterpret_cast<float4 *>(smem + SM_A2 + bsel * A2_BUF), *alo = ahi + (A2_BUF / 2) / 16;
#pragma unroll
            for (int kc = 0; kc < H2CHUNK / 4; ++kc) {
                ahi[kc * 128 + tid] = make_float4(hh[4 * kc], hh[4 * kc + 1], hh[4 * kc + 2], hh[4 * kc + 3]);
                alo[kc * 128 + tid] = make_float4(hl[4 * kc], hl[4 * kc + 1], hl[4 * kc + 2], hl[4 * kc + 3]);
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(full2 + bsel);   // -> the MMA warp issues this chunk's six MMAs into D2
        }
        // chunks 6 and 7 are still outstanding
        mbar_wait(bar2, ph2[0]);
        ph2[0] ^= 1;
        mbar_wait(bar2 + 1, ph2[1]);
        ph2[1] ^= 1;
        tc_fence_after();

        // ---- epilogue 2: bias, sigmoid, update mask, pre-update life mask, stores
        {
            float o[16];
            tmem_ld16(my_tmem + D2_COL, o);
            const size_t cell = ((size_t)b * a.H + yc) * a.W + xc;
            const bool upd = valid && nca_update_bit(a, cell);
            const float alpha_new = upd ? sigmoid_rn(o[3] + bias2[3]) : 0.f;
            const bool keep = upd && alive0;
            const uint32_t abits = __ballot_sync(0xffffffffu, valid && alpha_new > 0.1f);
            const uint32_t nbits = __ballot_sync(0xffffffffu, keep);
            if (y < a.H && lane == 0) {
                const size_t wi = ((size_t)b * a.H + y) * a.wpr + tx;
                a.alive_bits[wi] = abits;
                a.nz_bits[wi] = nbits;
            }
            if (valid) {
                float *ob = a.out + (size_t)b * NCA_N * plane + (size_t)y * a.W + x;
#pragma unroll
                for (int ch = 0; ch < NCA_N; ++ch) ob[(size_t)ch * plane] = keep ? sigmoid_rn(o[ch] + bias2[ch]) : 0.f;
            }
        }
        // all TMEM reads of this thread are done before it arrives on full1 for the next tile, and the next GEMM 1 is issued
        // only after all 128 arrivals: nothing overwrites D1 / D2 early
        tc_fence_before();
    }

    __syncthreads();
    if (threadIdx.x < 32) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(TMEM_COLS) : "memory");
    }
}

// ---- weight preparation: split hi/lo images in the operand layouts ----------------------------------------------
__global__ void nca_prepare_tc_kernel(const float *__restrict__ wp_simt, unsigned char *__restrict__ tc) {
    // wp_simt holds W1f[128][32], b1, W2t[128][16], b2 (see nca.cuh)
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    float *b1hi = reinterpret_cast<float *>(tc + TC_B1HI), *b1lo = reinterpret_cast<float *>(tc + TC_B1LO);
    float *b2hi = reinterpret_cast<float *>(tc + TC_B2HI), *b2lo = reinterpret_cast<float *>(tc + TC_B2LO);
    if (t < NCA_HID * K1P) {  // B1[n][k] = [W1f | b1 | 0][n][k] -> [k/4][n][k%4]
        const int n = t / K1P, k = t % K1P;
        const float w = k < NCA_K1 ? wp_simt[WP_W1F + n * NCA_K1 + k] : (k == NCA_K1 ? wp_simt[WP_B1 + n] : 0.f);
        const float hi = __uint_as_float((__float_as_uint(w) + 0x1000u) & 0xffffe000u);
        const int idx = ((k >> 2) * NCA_HID + n) * 4 + (k & 3);
        b1hi[idx] = hi;
        b1lo[idx] = w - hi;
    }
    if (t < NCA_HID * NCA_N) {  // B2[n][k] = W2[n][k] = W2t[k][n] -> [k/4][n][k%4]
        const int k = t / NCA_N, n = t % NCA_N;
        const float w = wp_simt[WP_W2T + t];
        const float hi = __uint_as_float((__float_as_uint(w) + 0x1000u) & 0xffffe000u);
        const int idx = ((k >> 2) * NCA_N + n) * 4 + (k & 3);
        b2hi[idx] = hi;
        b2lo[idx] = w - hi;
    }
    if (t < NCA_HID) reinterpret_cast<float *>(tc + TC_BIAS1)[t] = wp_simt[WP_B1 + t];
    if (t < NCA_N) reinterpret_cast<float *>(tc + TC_BIAS2)[t] = wp_simt[WP_B2 + t];
}

int nca_prepare_tc(const float *, const float *, const float *, const float *, void *Wprep, cudaStream_t st) {
    // runs after nca_prepare_simt_kernel on the same stream: reads the folded/transposed fp32 copies it produced
    unsigned char *base = reinterpret_cast<unsigned char *>(Wprep);
    nca_prepare_tc_kernel<<<ceil_div(NCA_HID * K1P, 256), 256, 0, st>>>(reinterpret_cast<const float *>(base), base + WP_TC_OFF);
    B200CA_LAUNCH_CHECK("nca_prepare_tc_kernel");
    return 0;
}

int nca_step_tc(const NcaArgs &a, cudaStream_t st) {
    static bool attr_done[64] = {false};
    int dev = 0;
    B200CA_CUDA(cudaGetDevice(&dev));
    if (!attr_done[dev & 63]) {
        B200CA_CUDA(cudaFuncSetAttribute(nca_update_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
        attr_done[dev & 63] = true;
    }
    B200CA_REQUIRE((long long)a.H * a.W <= (1LL << 30), "nca_step: a plane must stay below 2^30 cells (32-bit byte offsets)");
    const int tiles_x = ceil_div(a.W, 32), tiles_y = ceil_div(a.H, 4);
    const long long ntiles = (long long)tiles_x * tiles_y * a.B;
    const int grid = (int)std::min<long long>(ceil_div64(ntiles, NG), (long long)num_sms());
    B200CA_CUDA(launch_pdl(nca_update_tc_kernel, dim3(grid), dim3(NTHREADS), (size_t)SM_TOTAL, st, a, tiles_x, tiles_y));
    B200CA_LAUNCH_CHECK("nca_update_tc_kernel");
    note_variant("nca_update_tc<grid=%d>", grid);
    return 0;
}

}  // namespace b200ca
