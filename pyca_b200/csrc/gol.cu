// Bit-packed outer-totalistic 2-state CA (Game of Life family) for sm_100a.
//
// Replaces CA2D.step (pyca/automata/models/ca2d.py:195-209): the reference builds 8 rolled
// copies of an int32 world, adds them and selects bit `count` of the survive / birth mask
// (~22 elementwise passes, ~200 B/cell of traffic).  Here the world is 1 bit/cell:
//   * one thread owns VEC consecutive 32-cell words of a row and walks RPT rows downwards,
//     keeping a 3-row sliding window of bit-sliced horizontal sums in registers;
//   * west/east neighbours come from funnel shifts of (left word, word, right word); the left/
//     right words of a thread's group come from the adjacent lanes by warp shuffle, only the
//     warp-edge lanes touch memory for them;
//   * the 8-neighbour count is a bit-sliced adder tree (LOP3), the rule is evaluated on the
//     4 count bit-planes (2 LOP3 for Life, a mux tree for arbitrary 9-bit S/B masks).
// Algorithmic traffic: 1 bit read + 1 bit written per cell (0.25 B/cell/step).
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"

namespace b200ca {

// 32 cells of the periodic row starting at (possibly negative / >= W) cell index c0; generic slow path
// used only at the row ends when W is not a multiple of 32.
__device__ __noinline__ uint32_t gol_gather32(const uint32_t *__restrict__ row, int W, long long c0) {
    int c = (int)(((c0 % W) + W) % W);
    uint32_t out = 0;
    int filled = 0;
    while (filled < 32) {
        int wi = c >> 5, off = c & 31;
        int take = min(min(32 - off, W - c), 32 - filled);
        uint32_t bits = row[wi] >> off;
        if (take < 32) bits &= (1u << take) - 1u;
        out |= bits << filled;
        filled += take;
        c += take;
        if (c == W) c = 0;
    }
    return out;
}

struct GolGeom {
    int H, W, stride, wpr;  // wpr = words per row
    int aligned;            // W % 32 == 0
    uint32_t last_mask;     // valid bits of the last word of a row
    int vert_open;
};

// logical word j (j in [-1, wpr]) of row `row` of the x-periodic world
__device__ __forceinline__ uint32_t gol_fetch(const uint32_t *__restrict__ row, const GolGeom &g, int j) {
    if (g.aligned) {
        int jj = j < 0 ? j + g.wpr : (j >= g.wpr ? j - g.wpr : j);
        return __ldg(row + jj);
    }
    if (j >= 0 && j < g.wpr - 1) return __ldg(row + j);
    if (j >= g.wpr) return 0u;  // only feeds bits beyond W, which are masked off
    return gol_gather32(row, g.W, (long long)j * 32);
}

template <int VEC>
struct RowSums {
    uint32_t c[VEC];   // centre cells
    uint32_t s0[VEC], s1[VEC];  // west+centre+east (2-bit)
    uint32_t t0[VEC], t1[VEC];  // west+east (2-bit)
};

template <int VEC>
__device__ __forceinline__ void gol_load_row(RowSums<VEC> &R, const uint32_t *__restrict__ in, const GolGeom &g, int r,
                                             int j0, bool valid, int lane) {
    uint32_t c[VEC];
    bool live_row = true;
    int rr = r;
    if (g.vert_open) {
        live_row = (r >= 0 && r < g.H);
    } else {
        rr = r < 0 ? r + g.H : (r >= g.H ? r - g.H : r);
    }
    const uint32_t *row = in + (size_t)rr * g.stride;
    const bool ld = valid && live_row;
    if (VEC == 4) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (ld) v = __ldg(reinterpret_cast<const uint4 *>(row + j0));
        c[0] = v.x; c[1 % VEC] = v.y; c[2 % VEC] = v.z; c[3 % VEC] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) c[k] = ld ? gol_fetch(row, g, j0 + k) : 0u;
    }
    uint32_t left = __shfl_up_sync(0xffffffffu, c[VEC - 1], 1);
    uint32_t right = __shfl_down_sync(0xffffffffu, c[0], 1);
    if (lane == 0 || j0 == 0) left = ld ? gol_fetch(row, g, j0 - 1) : 0u;
    if (lane == 31 || j0 + VEC >= g.wpr) right = ld ? gol_fetch(row, g, j0 + VEC) : 0u;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        uint32_t L = k == 0 ? left : c[k - 1];
        uint32_t Rw = k == VEC - 1 ? right : c[k + 1];
        uint32_t w = __funnelshift_l(L, c[k], 1);   // bit i <- cell i-1
        uint32_t e = __funnelshift_r(c[k], Rw, 1);  // bit i <- cell i+1
        R.c[k] = c[k];
        R.t0[k] = w ^ e;
        R.t1[k] = w & e;
        R.s0[k] = R.t0[k] ^ c[k];
        R.s1[k] = R.t1[k] | (R.t0[k] & c[k]);
    }
}

struct RuleMasks {
    uint32_t s[9], b[9];  // 0 or ~0
};

template <bool LIFE>
__device__ __forceinline__ uint32_t gol_rule(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1, uint32_t m0, uint32_t m1,
                                             uint32_t c, const RuleMasks &rm) {
    // (a1a0) + (b1b0) -> x2x1x0 ; + (m1m0) -> y3y2y1y0
    uint32_t x0 = a0 ^ b0, k0 = a0 & b0;
    uint32_t x1 = a1 ^ b1 ^ k0;
    uint32_t x2 = (a1 & b1) | (k0 & (a1 ^ b1));
    uint32_t y0 = x0 ^ m0, d0 = x0 & m0;
    uint32_t y1 = x1 ^ m1 ^ d0;
    uint32_t d1 = (x1 & m1) | (d0 & (x1 ^ m1));
    uint32_t y2 = x2 ^ d1;
    if (LIFE) {
        // count==3 | (alive & count==2); count==8 has y1==0 so y3 is not needed
        return y1 & ~y2 & (y0 | c);
    } else {
        uint32_t y3 = x2 & d1;
        uint32_t m[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) m[k] = (c & rm.s[k]) | (~c & rm.b[k]);
        uint32_t q0 = (y0 & m[1]) | (~y0 & m[0]);
        uint32_t q1 = (y0 & m[3]) | (~y0 & m[2]);
        uint32_t q2 = (y0 & m[5]) | (~y0 & m[4]);
        uint32_t q3 = (y0 & m[7]) | (~y0 & m[6]);
        uint32_t h0 = (y1 & q1) | (~y1 & q0);
        uint32_t h1 = (y1 & q3) | (~y1 & q2);
        uint32_t f = (y2 & h1) | (~y2 & h0);
        return (y3 & m[8]) | (~y3 & f);
    }
}

template <int VEC, int RPT, bool LIFE>
__global__ void __launch_bounds__(128) gol_step_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, GolGeom g,
                                                        uint32_t s_mask, uint32_t b_mask) {
    const int lane = threadIdx.x & 31;
    const int unit = blockIdx.x * blockDim.x + threadIdx.x;
    const int j0 = unit * VEC;
    const bool valid = j0 < g.wpr;
    const int r0 = blockIdx.y * RPT;
    RuleMasks rm;
    if (!LIFE) {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            rm.s[k] = ((s_mask >> k) & 1u) ? 0xffffffffu : 0u;
            rm.b[k] = ((b_mask >> k) & 1u) ? 0xffffffffu : 0u;
        }
    }
    RowSums<VEC> A, Bc, Cn;
    gol_load_row<VEC>(A, in, g, r0 - 1, j0, valid, lane);
    gol_load_row<VEC>(Bc, in, g, r0, j0, valid, lane);
#pragma unroll 4
    for (int i = 0; i < RPT; ++i) {
        const int r = r0 + i;
        if (r >= g.H) break;  // uniform across the block
        gol_load_row<VEC>(Cn, in, g, r + 1, j0, valid, lane);
        uint32_t o[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            o[k] = gol_rule<LIFE>(A.s0[k], A.s1[k], Cn.s0[k], Cn.s1[k], Bc.t0[k], Bc.t1[k], Bc.c[k], rm);
            if (j0 + k == g.wpr - 1) o[k] &= g.last_mask;
        }
        if (valid) {
            uint32_t *orow = out + (size_t)r * g.stride + j0;
            if (VEC == 4) {
                *reinterpret_cast<uint4 *>(orow) = make_uint4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]);
            } else {
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    if (j0 + k < g.wpr) orow[k] = o[k];
            }
        }
        A = Bc;
        Bc = Cn;
    }
}

// ---- fast path: W % 128 == 0 (uint4 per thread), H % RPT == 0, no per-word fix-ups ---------------------
// Works on the 9-cell sums (centre included): per row and word only the 2-bit horizontal sum west+centre+east
// is kept (2 SHF + 2 LOP3, computed once per row, used by three output rows); an output word adds the three
// row sums bit-sliced (6 LOP3) and applies the rule on count9 (Life: count9==3 | alive & count9==4, 2 LOP3).
// Per row and thread additionally: 1 LDG.128, 1 predicated edge LDG.32 (warp-edge lanes), 2 SHFL, 1 STG.128.
// No branches in the row loop: threads past the row end load a clamped column and skip the store.
struct Row4 {
    uint32_t c[4], s0[4], s1[4];
};

// EVF: loads marked evict-first in L2.  The generation just read is dead (its buffer is overwritten two launches later) while the
// one being written is the next launch's input: when one buffer fits the 126 MB L2 -- a row slab of a sharded world -- the
// default policy streams both through the cache and keeps neither, evict-first reads leave the written generation resident.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <bool EVF>
__device__ __forceinline__ uint4 gol_ld4(const uint32_t *p, uint64_t pol) {
    if (!EVF) return __ldg(reinterpret_cast<const uint4 *>(p));
    uint4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}
template <bool EVF>
__device__ __forceinline__ uint32_t gol_ld1(const uint32_t *p, uint64_t pol) {
    if (!EVF) return __ldg(p);
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}

template <bool EVF = false>
__device__ __forceinline__ void gol_fast_load(Row4 &R, const uint32_t *__restrict__ base, uint32_t off_main, uint32_t off_edge,
                                              bool edge_l, bool edge_r, uint32_t live_mask, uint64_t pol = 0) {
    const uint4 v = gol_ld4<EVF>(base + off_main, pol);
    uint32_t ev = 0u;
    if (edge_l || edge_r) ev = gol_ld1<EVF>(base + off_edge, pol);
    uint32_t left = __shfl_up_sync(0xffffffffu, v.w, 1);
    uint32_t right = __shfl_down_sync(0xffffffffu, v.x, 1);
    left = edge_l ? ev : left;
    right = edge_r ? ev : right;
    const uint32_t c[4] = {v.x & live_mask, v.y & live_mask, v.z & live_mask, v.w & live_mask};
    left &= live_mask;
    right &= live_mask;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t L = k == 0 ? left : c[k - 1];
        const uint32_t Rw = k == 3 ? right : c[k + 1];
        const uint32_t w = __funnelshift_l(L, c[k], 1);
        const uint32_t e = __funnelshift_r(c[k], Rw, 1);
        R.c[k] = c[k];
        R.s0[k] = w ^ e ^ c[k];
        R.s1[k] = (w & e) | (c[k] & (w ^ e));
    }
}

// rule on the 9-cell count: p, q, n = 2-bit row sums (above, own, below), c = centre cells
template <bool LIFE>
__device__ __forceinline__ uint32_t gol_rule9(uint32_t p0, uint32_t p1, uint32_t q0, uint32_t q1, uint32_t n0, uint32_t n1, uint32_t c,
                                              const RuleMasks &rm) {
    const uint32_t z0 = p0 ^ q0 ^ n0;
    const uint32_t k0 = (p0 & q0) | (n0 & (p0 ^ q0));
    const uint32_t z1 = p1 ^ q1 ^ n1;
    const uint32_t k1 = (p1 & q1) | (n1 & (p1 ^ q1));
    const uint32_t y1 = k0 ^ z1;
    const uint32_t y2 = k1 ^ (k0 & z1);
    if (LIFE) {
        // count9 == 3 (z0=1,y1=1,y2=0)  |  alive & count9 == 4 (z0=0,y1=0,y2=1); count9 >= 8 has y1 == 0 and y2 == 0
        const uint32_t x = (y2 & c & ~z0) | (~y2 & z0);
        return x & ~(z0 ^ y1);
    } else {
        const uint32_t y3 = k1 & k0 & z1;
        uint32_t m[10];  // value for count9 == k: alive cells look at S[k-1], dead cells at B[k]
#pragma unroll
        for (int k = 0; k < 10; ++k) m[k] = (k > 0 ? (c & rm.s[k - 1]) : 0u) | (k < 9 ? (~c & rm.b[k]) : 0u);
        const uint32_t q_0 = (z0 & m[1]) | (~z0 & m[0]);
        const uint32_t q_1 = (z0 & m[3]) | (~z0 & m[2]);
        const uint32_t q_2 = (z0 & m[5]) | (~z0 & m[4]);
        const uint32_t q_3 = (z0 & m[7]) | (~z0 & m[6]);
        const uint32_t q_4 = (z0 & m[9]) | (~z0 & m[8]);
        const uint32_t h0 = (y1 & q_1) | (~y1 & q_0);
        const uint32_t h1 = (y1 & q_3) | (~y1 & q_2);
        const uint32_t f = (y2 & h1) | (~y2 & h0);
        return (y3 & q_4) | (~y3 & f);
    }
}

template <bool LIFE, bool OPEN, int RPT, bool EVF>
__global__ void __launch_bounds__(128, 6) gol_fast_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int H, int wpr,
                                                           int stride, uint32_t s_mask, uint32_t b_mask) {
    const int lane = threadIdx.x & 31;
    const int units = wpr >> 2;
    const int unit_raw = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = unit_raw < units;
    const int unit = valid ? unit_raw : units - 1;
    const uint32_t j0 = unit * 4;
    // warp-edge lanes fetch the neighbouring word from memory; (units % 32 != 1 is guaranteed by the launcher, so
    // no lane needs both)
    const bool edge_l = (lane == 0), edge_r = !edge_l && ((lane == 31) || (unit == units - 1));
    const uint32_t je = edge_l ? (unit == 0 ? wpr : j0) - 1 : ((unit == units - 1) ? 0 : j0 + 4);
    const int r0 = blockIdx.y * RPT;  // H % RPT == 0: every block owns RPT full rows
    RuleMasks rm;
    if (!LIFE) {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            rm.s[k] = ((s_mask >> k) & 1u) ? 0xffffffffu : 0u;
            rm.b[k] = ((b_mask >> k) & 1u) ? 0xffffffffu : 0u;
        }
    }
    const uint64_t pol = EVF ? l2_evict_first_policy() : 0;
    pdl_wait_then_release();  // generations are chained with programmatic dependent launch
    // rows r0 .. r0+RPT-1 are contiguous; only the row above the first and below the last may wrap / be dead
    const int ra = r0 == 0 ? (OPEN ? 0 : H - 1) : r0 - 1;
    const int rb = r0 + RPT == H ? (OPEN ? H - 1 : 0) : r0 + RPT;
    const uint32_t ma = (OPEN && r0 == 0) ? 0u : 0xffffffffu;
    const uint32_t mb = (OPEN && r0 + RPT == H) ? 0u : 0xffffffffu;
    const uint32_t ustride = (uint32_t)stride;
    uint32_t row_off = (uint32_t)r0 * ustride;  // H * stride < 2^32 words is guaranteed by the launcher
    Row4 Q, N;
    uint32_t a0[4], a1[4];
    gol_fast_load<EVF>(N, in, (uint32_t)ra * ustride + j0, (uint32_t)ra * ustride + je, edge_l, edge_r, ma, pol);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a0[k] = N.s0[k];
        a1[k] = N.s1[k];
    }
    gol_fast_load<EVF>(Q, in, row_off + j0, row_off + je, edge_l, edge_r, 0xffffffffu, pol);
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const uint32_t next_off = row_off + ustride;
        if (i == RPT - 1)
            gol_fast_load<EVF>(N, in, (uint32_t)rb * ustride + j0, (uint32_t)rb * ustride + je, edge_l, edge_r, mb, pol);
        else
            gol_fast_load<EVF>(N, in, next_off + j0, next_off + je, edge_l, edge_r, 0xffffffffu, pol);
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o[k] = gol_rule9<LIFE>(a0[k], a1[k], Q.s0[k], Q.s1[k], N.s0[k], N.s1[k], Q.c[k], rm);
            a0[k] = Q.s0[k];
            a1[k] = Q.s1[k];
        }
        if (valid) *reinterpret_cast<uint4 *>(out + (row_off + j0)) = make_uint4(o[0], o[1], o[2], o[3]);
        row_off = next_off;
        Q = N;
    }
}

// ---- small worlds: the whole torus lives in shared memory, nsteps generations per launch ---------------------------
// A 250x250 world (simulate.py's default, BASELINE configs[0]) is 2000 words: a launch per generation is pure launch
// latency (~5-10 us for ~0.1 us of work).  One CTA keeps both ping-pong copies in shared memory, runs all the
// generations with a __syncthreads() between them and writes the result once.  Any width (partial last word) and
// both vertical modes; used by b200ca_gol_run when H * words_per_row <= GOL_SMALL_WORDS (beyond that one SM is
// slower than a launch per generation on 148 SMs).
constexpr int GOL_SMALL_THREADS = 1024;
constexpr int GOL_SMALL_WPT = 4;                                   // words per thread
constexpr int GOL_SMALL_WORDS = GOL_SMALL_THREADS * GOL_SMALL_WPT;  // 4096 words = 131072 cells

// Word j (j in [-1, wpr]) of a row as far as a 1-cell horizontal shift needs it: the wrap of a row whose last word is
// partial (nb valid bits) is the last word moved up so cell W-1 sits in bit 31 (west of cell 0), resp. cell 0 injected
// after cell W-1 (east of cell W-1).  Branch-light replacement of the generic bit gather.
__device__ __forceinline__ uint32_t gol_fetch_any(const uint32_t *row, const GolGeom &g, int nb, int j) {
    if (j >= 0 && j < g.wpr - 1) return row[j];
    if (j == g.wpr - 1) return g.aligned ? row[j] : (row[j] | (row[0] << nb));
    if (j < 0) return g.aligned ? row[g.wpr - 1] : (row[g.wpr - 1] << (32 - nb));
    return g.aligned ? row[0] : 0u;
}

// 2-bit sum west+centre+east of word j of `row` (zeros when the row is outside an open world)
__device__ __forceinline__ void gol_row_sum(const uint32_t *row, uint32_t live_mask, const GolGeom &g, int nb, int j, uint32_t &c,
                                            uint32_t &s0, uint32_t &s1) {
    const uint32_t C = gol_fetch_any(row, g, nb, j) & live_mask;
    const uint32_t L = gol_fetch_any(row, g, nb, j - 1) & live_mask;
    const uint32_t Rw = gol_fetch_any(row, g, nb, j + 1) & live_mask;
    const uint32_t w = __funnelshift_l(L, C, 1), e = __funnelshift_r(C, Rw, 1);
    c = C;
    s0 = w ^ e ^ C;
    s1 = (w & e) | (C & (w ^ e));
}

template <bool LIFE>
__global__ void __launch_bounds__(GOL_SMALL_THREADS) gol_small_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, GolGeom g,
                                                                        uint32_t s_mask, uint32_t b_mask, int nsteps) {
    extern __shared__ uint32_t gsm[];
    const int total = g.H * g.wpr;
    const int nb = g.W - 32 * (g.wpr - 1);
    uint32_t *cur = gsm, *nxt = gsm + total;
    RuleMasks rm;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        rm.s[k] = ((s_mask >> k) & 1u) ? 0xffffffffu : 0u;
        rm.b[k] = ((b_mask >> k) & 1u) ? 0xffffffffu : 0u;
    }
    // this thread's words and their row neighbourhood, fixed for all generations
    int off_a[GOL_SMALL_WPT], off_q[GOL_SMALL_WPT], off_b[GOL_SMALL_WPT], jj[GOL_SMALL_WPT];
    uint32_t ma[GOL_SMALL_WPT], mb[GOL_SMALL_WPT];
#pragma unroll
    for (int k = 0; k < GOL_SMALL_WPT; ++k) {
        const int i = threadIdx.x + k * GOL_SMALL_THREADS;
        const int r = i < total ? i / g.wpr : 0;
        jj[k] = i < total ? i - r * g.wpr : -100;  // -100: no word
        int ra = r - 1, rb = r + 1;
        ma[k] = mb[k] = 0xffffffffu;
        if (g.vert_open) {
            if (ra < 0) { ra = 0; ma[k] = 0u; }
            if (rb >= g.H) { rb = g.H - 1; mb[k] = 0u; }
        } else {
            ra = ra < 0 ? g.H - 1 : ra;
            rb = rb >= g.H ? 0 : rb;
        }
        off_a[k] = ra * g.wpr;
        off_q[k] = r * g.wpr;
        off_b[k] = rb * g.wpr;
        if (i < total) cur[i] = in[(size_t)r * g.stride + jj[k]];
    }
    __syncthreads();
    for (int step = 0; step < nsteps; ++step) {
#pragma unroll
        for (int k = 0; k < GOL_SMALL_WPT; ++k) {
            if (jj[k] < 0) continue;
            const int j = jj[k];
            uint32_t ca, a0, a1, cq, q0, q1, cb, b0, b1;
            gol_row_sum(cur + off_a[k], ma[k], g, nb, j, ca, a0, a1);
            gol_row_sum(cur + off_q[k], 0xffffffffu, g, nb, j, cq, q0, q1);
            gol_row_sum(cur + off_b[k], mb[k], g, nb, j, cb, b0, b1);
            uint32_t o = gol_rule9<LIFE>(a0, a1, q0, q1, b0, b1, cq, rm);
            if (j == g.wpr - 1) o &= g.last_mask;
            nxt[off_q[k] + j] = o;
        }
        __syncthreads();
        uint32_t *t = cur;
        cur = nxt;
        nxt = t;
    }
#pragma unroll
    for (int k = 0; k < GOL_SMALL_WPT; ++k)
        if (jj[k] >= 0) out[(size_t)(off_q[k] / g.wpr) * g.stride + jj[k]] = cur[off_q[k] + jj[k]];
}

// ---- tiny worlds (<= 256 cells wide, <= 1024 rows): one ROW per thread, held in registers for all generations ----------
// The 250x250 world of simulate.py is 8 words by 250 rows.  A thread owns its whole row, so the horizontal neighbourhood
// (with the periodic wrap through the partial last word) never leaves its registers; per generation it publishes the row's
// 2-bit sums west+centre+east as 8 uint2 in shared memory, one __syncthreads() later reads the two neighbouring rows' sums,
// and applies the rule on the 9-cell count.  The exchange buffer ping-pongs by generation parity, so one barrier per
// generation is enough.  ~170 instructions per thread and generation instead of nine shared-memory row fetches with wrap
// logic per WORD in gol_small_kernel.
template <int WPR, bool LIFE>
__global__ void __launch_bounds__(1024) gol_rows_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, GolGeom g,
                                                        uint32_t s_mask, uint32_t b_mask, int nsteps) {
    extern __shared__ uint2 grs[];  // [2][WPR][rows_pad]
    const int r = threadIdx.x, rows_pad = blockDim.x;
    const bool live = r < g.H;
    const int nb = g.W - 32 * (WPR - 1);  // valid bits of the last word
    RuleMasks rm;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        rm.s[k] = ((s_mask >> k) & 1u) ? 0xffffffffu : 0u;
        rm.b[k] = ((b_mask >> k) & 1u) ? 0xffffffffu : 0u;
    }
    uint32_t w[WPR];
#pragma unroll
    for (int j = 0; j < WPR; ++j) w[j] = live ? in[(size_t)r * g.stride + j] : 0u;
    w[WPR - 1] &= g.last_mask;
    // rows above / below (periodic, or nothing beyond the edge of an open world)
    const int ra = r > 0 ? r - 1 : g.H - 1, rb = r + 1 < g.H ? r + 1 : 0;
    const uint32_t ma = (g.vert_open && r == 0) ? 0u : 0xffffffffu, mb = (g.vert_open && r == g.H - 1) ? 0u : 0xffffffffu;
    for (int step = 0; step < nsteps; ++step) {
        uint2 *buf = grs + (size_t)(step & 1) * WPR * rows_pad;
        uint32_t q0[WPR], q1[WPR];
#pragma unroll
        for (int j = 0; j < WPR; ++j) {
            // west: cell x-1 at bit x; the cell left of x = 0 is cell W-1 = bit nb-1 of the last word (and the other way round)
            const uint32_t prev = j > 0 ? w[j - 1] >> 31 : (w[WPR - 1] >> (nb - 1)) & 1u;
            const uint32_t west = (w[j] << 1) | prev;
            uint32_t east = w[j] >> 1;
            if (j < WPR - 1) east |= w[j + 1] << 31;
            else east |= (w[0] & 1u) << (nb - 1);
            q0[j] = west ^ east ^ w[j];
            q1[j] = (west & east) | (w[j] & (west ^ east));
            if (j == WPR - 1) {  // the shifted-in bits beyond the world stay out of the sums
                q0[j] &= g.last_mask;
                q1[j] &= g.last_mask;
            }
            buf[j * rows_pad + r] = make_uint2(q0[j], q1[j]);
        }
        __syncthreads();
        if (live) {
#pragma unroll
            for (int j = 0; j < WPR; ++j) {
                const uint2 a = buf[j * rows_pad + ra], b = buf[j * rows_pad + rb];
                w[j] = gol_rule9<LIFE>(a.x & ma, a.y & ma, q0[j], q1[j], b.x & mb, b.y & mb, w[j], rm);
            }
            w[WPR - 1] &= g.last_mask;
        }
    }
    if (live) {
#pragma unroll
        for (int j = 0; j < WPR; ++j) out[(size_t)r * g.stride + j] = w[j];
    }
}

template <int WPR>
static int launch_gol_rows(const uint32_t *a, uint32_t *dst, const GolGeom &g, uint32_t s_mask, uint32_t b_mask, int nsteps,
                           cudaStream_t st) {
    const int threads = (g.H + 31) / 32 * 32;
    const size_t smem = (size_t)2 * WPR * threads * sizeof(uint2);
    const bool life = (s_mask == 0b1100u && b_mask == 0b1000u);
    if (life) {
        if (smem > 48 * 1024) B200CA_CUDA(cudaFuncSetAttribute(gol_rows_kernel<WPR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gol_rows_kernel<WPR, true><<<1, threads, smem, st>>>(a, dst, g, s_mask, b_mask, nsteps);
        note_variant("gol_rows<%d,life>", WPR);
    } else {
        if (smem > 48 * 1024) B200CA_CUDA(cudaFuncSetAttribute(gol_rows_kernel<WPR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gol_rows_kernel<WPR, false><<<1, threads, smem, st>>>(a, dst, g, s_mask, b_mask, nsteps);
        note_variant("gol_rows<%d,generic>", WPR);
    }
    B200CA_LAUNCH_CHECK("gol_rows_kernel");
    return 0;
}

// ---- pack / unpack / population / random fill ---------------------------------------------------
template <typename T>
__global__ void gol_pack_kernel(const T *__restrict__ world, uint32_t *__restrict__ bits, int H, int W, int stride, int wpr) {
    // one warp packs 32 consecutive words of a row (1024 cells): 32 coalesced reads + ballots
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int groups = ceil_div(wpr, 32);
    const int y = warp / groups;
    if (y >= H) return;
    const int wg = (warp % groups) * 32;
    const T *row = world + (size_t)y * W;
    uint32_t mine = 0;
    for (int k = 0; k < 32; ++k) {
        int x = (wg + k) * 32 + lane;
        bool alive = (x < W) && (row[x] == (T)1);
        uint32_t b = __ballot_sync(0xffffffffu, alive);
        if (k == lane) mine = b;
    }
    if (wg + lane < wpr) bits[(size_t)y * stride + wg + lane] = mine;
}

template <typename T>
__global__ void gol_unpack_kernel(const uint32_t *__restrict__ bits, T *__restrict__ world, int H, int W, int stride) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    uint32_t wv = __ldg(bits + (size_t)y * stride + (x >> 5));
    world[(size_t)y * W + x] = (T)((wv >> (x & 31)) & 1u);
}

__global__ void gol_population_kernel(const uint32_t *__restrict__ bits, int H, int wpr, int stride,
                                      unsigned long long *__restrict__ count) {
    unsigned long long local = 0;
    const size_t total = (size_t)H * wpr;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t y = i / wpr, j = i % wpr;
        local += __popc(bits[y * stride + j]);
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ unsigned long long sm[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sm[wid] = local;
    __syncthreads();
    if (wid == 0) {
        local = lane < (blockDim.x >> 5) ? sm[lane] : 0ull;
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if (lane == 0) atomicAdd(count, local);
    }
}

__global__ void gol_random_fill_kernel(uint32_t *__restrict__ bits, int H, int wpr, int stride, uint32_t last_mask,
                                       uint64_t seed, int64_t row0) {
    // one Philox call -> 4 words
    const int quads = ceil_div(wpr, 4);
    const size_t total = (size_t)H * quads;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t y = i / quads;
        int q = (int)(i % quads);
        uint32_t r[4];
        Philox::gen(seed, (uint64_t)q, (uint64_t)(row0 + (int64_t)y), r);
        for (int k = 0; k < 4; ++k) {
            int j = q * 4 + k;
            if (j < wpr) {
                uint32_t v = r[k];
                if (j == wpr - 1) v &= last_mask;
                bits[y * stride + j] = v;
            }
        }
    }
}

static int make_geom(GolGeom &g, int H, int W, int stride, int vert_open) {
    B200CA_REQUIRE(H >= 1 && W >= 1, "gol: bad world size %dx%d", H, W);
    g.H = H;
    g.W = W;
    g.wpr = ceil_div(W, 32);
    B200CA_REQUIRE(stride >= g.wpr, "gol: stride_words %d < words per row %d", stride, g.wpr);
    g.stride = stride;
    g.aligned = (W % 32 == 0);
    int nb = W - 32 * (g.wpr - 1);
    g.last_mask = nb == 32 ? 0xffffffffu : ((1u << nb) - 1u);
    g.vert_open = vert_open ? 1 : 0;
    return 0;
}

template <int VEC, int RPT>
static int launch_step(const uint32_t *in, uint32_t *out, const GolGeom &g, uint32_t s, uint32_t b, cudaStream_t st) {
    const int units = ceil_div(g.wpr, VEC);
    int bx = units >= 128 ? 128 : ceil_div(units, 32) * 32;
    dim3 grid(ceil_div(units, bx), ceil_div(g.H, RPT));
    if (s == 0b1100u && b == 0b1000u)
        gol_step_kernel<VEC, RPT, true><<<grid, bx, 0, st>>>(in, out, g, s, b);
    else
        gol_step_kernel<VEC, RPT, false><<<grid, bx, 0, st>>>(in, out, g, s, b);
    B200CA_LAUNCH_CHECK("gol_step_kernel");
    return 0;
}

template <bool LIFE, bool OPEN, int RPT>
static int launch_fast_t(const uint32_t *in, uint32_t *out, const GolGeom &g, uint32_t s, uint32_t b, cudaStream_t st) {
    const int units = g.wpr / 4;
    const int bx = units >= 128 ? 128 : ceil_div(units, 32) * 32;
    dim3 grid(ceil_div(units, bx), ceil_div(g.H, RPT));
    // one buffer fits L2 with room to spare (but the pair does not anyway fit trivially): read the dying generation evict-first
    const long long bytes = (long long)g.H * g.stride * 4;
    bool evf = bytes <= (100LL << 20) && bytes >= (24LL << 20);
#ifdef B200CA_DEV
    if (const char *e = getenv("B200CA_GOL_EVF")) evf = atoi(e) != 0;
#endif
    if (evf) B200CA_CUDA(launch_pdl(gol_fast_kernel<LIFE, OPEN, RPT, true>, grid, dim3(bx), 0, st, in, out, g.H, g.wpr, g.stride, s, b));
    else B200CA_CUDA(launch_pdl(gol_fast_kernel<LIFE, OPEN, RPT, false>, grid, dim3(bx), 0, st, in, out, g.H, g.wpr, g.stride, s, b));
    note_variant("gol_fast<%d,%d,%d%s>", (int)LIFE, (int)OPEN, RPT, evf ? ",evict-first" : "");
    return 0;
}

static int launch_fast(const uint32_t *in, uint32_t *out, const GolGeom &g, uint32_t s, uint32_t b, cudaStream_t st) {
    const bool life = (s == 0b1100u && b == 0b1000u);
    // Rows per thread: 16 amortises the two halo rows best, but a slab of a sharded world may then be only 2-3 waves of
    // blocks and lose a quarter of the machine to the last partial wave: take the largest of {16, 8, 4} that divides H
    // and keeps the wave quantisation loss under 10 % (6 resident blocks of 128 threads per SM).
    const long long slots = (long long)num_sms() * 6;
    int rpt = 4;
    double best = -1.0;
    const int cand[3] = {16, 8, 4};
    for (int c = 0; c < 3; ++c) {
        if (g.H % cand[c]) continue;
        const long long blocks = (long long)ceil_div(g.wpr / 4, 128) * (g.H / cand[c]);
        const double eff = (double)blocks / (double)(ceil_div64(blocks, slots) * slots);
        const double score = eff >= 0.9 ? 2.0 + cand[c] : eff;   // first (largest) candidate above 90 % wins
        if (score > best) {
            best = score;
            rpt = cand[c];
        }
    }
#ifdef B200CA_DEV
    if (const char *e = getenv("B200CA_GOL_RPT")) rpt = atoi(e) == 16 ? 16 : (atoi(e) == 8 ? 8 : 4);
#endif
#define B200CA_GOL_DISPATCH(LIFE_, OPEN_)                                                        \
    (rpt == 16 ? launch_fast_t<LIFE_, OPEN_, 16>(in, out, g, s, b, st)                           \
               : (rpt == 8 ? launch_fast_t<LIFE_, OPEN_, 8>(in, out, g, s, b, st) : launch_fast_t<LIFE_, OPEN_, 4>(in, out, g, s, b, st)))
    if (life) return g.vert_open ? B200CA_GOL_DISPATCH(true, true) : B200CA_GOL_DISPATCH(true, false);
    return g.vert_open ? B200CA_GOL_DISPATCH(false, true) : B200CA_GOL_DISPATCH(false, false);
#undef B200CA_GOL_DISPATCH
}

static int gol_step_impl(const uint32_t *in, uint32_t *out, int H, int W, int stride, uint32_t s, uint32_t b, int vert_open,
                         cudaStream_t st) {
    GolGeom g;
    int rc = make_geom(g, H, W, stride, vert_open);
    if (rc) return rc;
    B200CA_REQUIRE(in && out && in != out, "gol_step: null or aliased buffers");
    B200CA_REQUIRE(s < 512 && b < 512, "gol_step: rule masks are 9-bit");
    const bool vec4 = g.aligned && (g.wpr % 4 == 0) && (stride % 4 == 0) && ((((uintptr_t)in) | ((uintptr_t)out)) % 16 == 0);
    if (vec4 && (g.H % 4 == 0) && ((g.wpr / 4) % 32 != 1 || g.wpr == 4) && (long long)g.H * g.stride < (1ll << 32) && g.wpr > 4)
        return launch_fast(in, out, g, s, b, st);
    // enough row bands to fill the machine: 16 rows per thread for big worlds, 4 for small ones
    const long long units = (long long)ceil_div(g.wpr, vec4 ? 4 : 1) * ceil_div(H, 16);
    const bool big = units >= (long long)num_sms() * 128 * 8;
    if (vec4) return big ? launch_step<4, 16>(in, out, g, s, b, st) : launch_step<4, 4>(in, out, g, s, b, st);
    return big ? launch_step<1, 16>(in, out, g, s, b, st) : launch_step<1, 4>(in, out, g, s, b, st);
}

}  // namespace b200ca

using namespace b200ca;

extern "C" int b200ca_gol_words_per_row(int W) { return ceil_div(W, 32); }

extern "C" int b200ca_gol_pack(const void *world, int elem_bytes, uint32_t *bits, int H, int W, int stride_words, void *stream) {
    GolGeom g;
    int rc = make_geom(g, H, W, stride_words, 0);
    if (rc) return rc;
    B200CA_REQUIRE(world && bits, "gol_pack: null pointer");
    const int groups = ceil_div(g.wpr, 32);
    const long long warps = (long long)H * groups;
    const int threads = 256;
    const unsigned blocks = (unsigned)ceil_div64(warps * 32, threads);
    cudaStream_t st = as_stream(stream);
    switch (elem_bytes) {
        case 1: gol_pack_kernel<uint8_t><<<blocks, threads, 0, st>>>((const uint8_t *)world, bits, H, W, stride_words, g.wpr); break;
        case 2: gol_pack_kernel<int16_t><<<blocks, threads, 0, st>>>((const int16_t *)world, bits, H, W, stride_words, g.wpr); break;
        case 4: gol_pack_kernel<int32_t><<<blocks, threads, 0, st>>>((const int32_t *)world, bits, H, W, stride_words, g.wpr); break;
        case 8: gol_pack_kernel<long long><<<blocks, threads, 0, st>>>((const long long *)world, bits, H, W, stride_words, g.wpr); break;
        default: B200CA_REQUIRE(false, "gol_pack: elem_bytes %d not in {1,2,4,8}", elem_bytes);
    }
    B200CA_LAUNCH_CHECK("gol_pack_kernel");
    return 0;
}

extern "C" int b200ca_gol_unpack(const uint32_t *bits, void *world, int elem_bytes, int H, int W, int stride_words, void *stream) {
    GolGeom g;
    int rc = make_geom(g, H, W, stride_words, 0);
    if (rc) return rc;
    B200CA_REQUIRE(world && bits, "gol_unpack: null pointer");
    dim3 grid(ceil_div(W, 256), H);
    cudaStream_t st = as_stream(stream);
    switch (elem_bytes) {
        case 1: gol_unpack_kernel<uint8_t><<<grid, 256, 0, st>>>(bits, (uint8_t *)world, H, W, stride_words); break;
        case 2: gol_unpack_kernel<int16_t><<<grid, 256, 0, st>>>(bits, (int16_t *)world, H, W, stride_words); break;
        case 4: gol_unpack_kernel<int32_t><<<grid, 256, 0, st>>>(bits, (int32_t *)world, H, W, stride_words); break;
        case 8: gol_unpack_kernel<long long><<<grid, 256, 0, st>>>(bits, (long long *)world, H, W, stride_words); break;
        default: B200CA_REQUIRE(false, "gol_unpack: elem_bytes %d not in {1,2,4,8}", elem_bytes);
    }
    B200CA_LAUNCH_CHECK("gol_unpack_kernel");
    return 0;
}

extern "C" int b200ca_gol_step(const uint32_t *bits_in, uint32_t *bits_out, int H, int W, int stride_words, uint32_t s_mask,
                               uint32_t b_mask, int vert_open, void *stream) {
    return gol_step_impl(bits_in, bits_out, H, W, stride_words, s_mask, b_mask, vert_open, as_stream(stream));
}

extern "C" int b200ca_gol_run(uint32_t *a, uint32_t *b, int H, int W, int stride_words, uint32_t s_mask, uint32_t b_mask,
                              int vert_open, int nsteps, void *stream) {
    B200CA_REQUIRE(nsteps >= 0, "gol_run: nsteps < 0");
    {
        // small worlds: all generations in one launch, world resident in shared memory; result lands where the
        // ping-pong contract says (a for even nsteps, b for odd)
        GolGeom g;
        int rc = make_geom(g, H, W, stride_words, vert_open);
        if (rc) return rc;
        if (nsteps >= 2 && (long long)H * g.wpr <= GOL_SMALL_WORDS) {
            B200CA_REQUIRE(a && b && a != b, "gol_run: null or aliased buffers");
            B200CA_REQUIRE(s_mask < 512 && b_mask < 512, "gol_run: rule masks are 9-bit");
            const size_t smem = (size_t)2 * H * g.wpr * sizeof(uint32_t);
            const bool life = (s_mask == 0b1100u && b_mask == 0b1000u);
            uint32_t *dst = (nsteps & 1) ? b : a;
            cudaStream_t st = as_stream(stream);
            if (g.wpr <= 8 && H <= 1024 && H >= 2) {  // a row per thread, in registers
                switch (g.wpr) {
                    case 1: return launch_gol_rows<1>(a, dst, g, s_mask, b_mask, nsteps, st);
                    case 2: return launch_gol_rows<2>(a, dst, g, s_mask, b_mask, nsteps, st);
                    case 3: return launch_gol_rows<3>(a, dst, g, s_mask, b_mask, nsteps, st);
                    case 4: return launch_gol_rows<4>(a, dst, g, s_mask, b_mask, nsteps, st);
                    case 5: return launch_gol_rows<5>(a, dst, g, s_mask, b_mask, nsteps, st);
                    case 6: return launch_gol_rows<6>(a, dst, g, s_mask, b_mask, nsteps, st);
                    case 7: return launch_gol_rows<7>(a, dst, g, s_mask, b_mask, nsteps, st);
                    default: return launch_gol_rows<8>(a, dst, g, s_mask, b_mask, nsteps, st);
                }
            }
            if (life) {
                if (smem > 48 * 1024) B200CA_CUDA(cudaFuncSetAttribute(gol_small_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                gol_small_kernel<true><<<1, GOL_SMALL_THREADS, smem, st>>>(a, dst, g, s_mask, b_mask, nsteps);
                note_variant("gol_small<life>");
            } else {
                if (smem > 48 * 1024) B200CA_CUDA(cudaFuncSetAttribute(gol_small_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                gol_small_kernel<false><<<1, GOL_SMALL_THREADS, smem, st>>>(a, dst, g, s_mask, b_mask, nsteps);
                note_variant("gol_small<generic>");
            }
            B200CA_LAUNCH_CHECK("gol_small_kernel");
            return 0;
        }
    }
    uint32_t *src = a, *dst = b;
    for (int i = 0; i < nsteps; ++i) {
        int rc = gol_step_impl(src, dst, H, W, stride_words, s_mask, b_mask, vert_open, as_stream(stream));
        if (rc) return rc;
        uint32_t *t = src;
        src = dst;
        dst = t;
    }
    return 0;
}

extern "C" int b200ca_gol_population(const uint32_t *bits, int H, int W, int stride_words, unsigned long long *count, void *stream) {
    GolGeom g;
    int rc = make_geom(g, H, W, stride_words, 0);
    if (rc) return rc;
    B200CA_REQUIRE(bits && count, "gol_population: null pointer");
    cudaStream_t st = as_stream(stream);
    B200CA_CUDA(cudaMemsetAsync(count, 0, sizeof(unsigned long long), st));
    const long long total = (long long)H * g.wpr;
    int blocks = (int)std::min<long long>(ceil_div64(total, 256), (long long)num_sms() * 8);
    gol_population_kernel<<<blocks, 256, 0, st>>>(bits, H, g.wpr, stride_words, count);
    B200CA_LAUNCH_CHECK("gol_population_kernel");
    return 0;
}

extern "C" int b200ca_gol_random_fill(uint32_t *bits, int H, int W, int stride_words, uint64_t seed, int64_t row0, void *stream) {
    GolGeom g;
    int rc = make_geom(g, H, W, stride_words, 0);
    if (rc) return rc;
    B200CA_REQUIRE(bits, "gol_random_fill: null pointer");
    const long long total = (long long)H * ceil_div(g.wpr, 4);
    int blocks = (int)std::min<long long>(ceil_div64(total, 256), (long long)num_sms() * 16);
    gol_random_fill_kernel<<<blocks, 256, 0, as_stream(stream)>>>(bits, H, g.wpr, stride_words, g.last_mask, seed, row0);
    B200CA_LAUNCH_CHECK("gol_random_fill_kernel");
    return 0;
}

extern "C" int b200ca_gol_run_host(const int32_t *h_world_in, int32_t *h_world_out, int H, int W, uint32_t s_mask,
                                   uint32_t b_mask, int nsteps) {
    B200CA_REQUIRE(h_world_in && h_world_out && H >= 1 && W >= 1 && nsteps >= 0, "gol_run_host: bad arguments");
    const int wpr = ceil_div(W, 32);
    const int stride = ceil_div(wpr, 4) * 4;
    int32_t *d_world = nullptr;
    uint32_t *d_a = nullptr, *d_b = nullptr;
    const size_t wbytes = (size_t)H * W * sizeof(int32_t), bbytes = (size_t)H * stride * sizeof(uint32_t);
    int rc = 0;
    cudaStream_t st = 0;
    do {
        if ((rc = scratch_get(0, wbytes, (void **)&d_world))) break;
        if ((rc = scratch_get(1, bbytes, (void **)&d_a))) break;
        if ((rc = scratch_get(2, bbytes, (void **)&d_b))) break;
        if ((rc = check_cuda(cudaMemcpyAsync(d_world, h_world_in, wbytes, cudaMemcpyHostToDevice, st), "H2D"))) break;
        if ((rc = b200ca_gol_pack(d_world, 4, d_a, H, W, stride, st))) break;
        if ((rc = b200ca_gol_run(d_a, d_b, H, W, stride, s_mask, b_mask, 0, nsteps, st))) break;
        if ((rc = b200ca_gol_unpack((nsteps & 1) ? d_b : d_a, d_world, 4, H, W, stride, st))) break;
        if ((rc = check_cuda(cudaMemcpyAsync(h_world_out, d_world, wbytes, cudaMemcpyDeviceToHost, st), "D2H"))) break;
        rc = check_cuda(cudaStreamSynchronize(st), "sync");
    } while (0);
    return rc;
}
