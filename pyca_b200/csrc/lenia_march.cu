// Lenia-family step, "marching" FFT path for sm_100a: ONE kernel per source channel, spectra never leave the SM.
//
// Same decomposition as lenia_fft.cu -- row FFT x real 31-tap column FIR on the spectra x inverse row FFT, fused with the
// growth function, the weighted source sum and the Euler step / clamp (lenia.py:430-469; affinity only for the MaCE
// family, macelenia.py:122-159) -- but organised so that the row spectra are produced, filtered and consumed inside one
// CTA's shared memory instead of round-tripping through HBM/L2 between a row-FFT kernel and a FIR+inverse kernel:
//
//   * x direction: overlap-save strips of N = 256 samples, V = 224 valid outputs (16-sample halo on both sides, so the
//     strip starts on a 64-byte boundary of the framed row and the frame's 16 periodic columns are exactly the halo).
//   * y direction: a CTA marches down a BAND of 2*Lp image rows of one strip, 8 row pairs at a time.  Rows y0+p and y0+Lp+p
//     ride in one complex transform (real taps keep re and im apart), so a band needs Lp+30 forward transforms for 2*Lp
//     output rows.  The last 40 spectra rows live in a shared-memory ring (5 blocks of 8 rows).
//   * which bands a CTA gets: a whole-world / whole-slab step is a BALANCED launch -- one wave of CTAs (2 per SM) shares the
//     columns of bands of all worlds and strips in equal-cost runs, or with every column cut into k bands (next_band() below,
//     b200ca_lenia_march_plan* is its host copy); a launch restricted to a band range keeps fixed-height bands (choose_lp).
//   * per iteration: FIR of 8 output row pairs from 38 ring rows (FFMA2 on (re0,re1)/(im0,im1) pairs, all 256 threads) ->
//     [warps 0-3: 8 inverse transforms | warps 4-7: 8 forward transforms of the block needed 5 iterations later] ->
//     epilogue on all threads (growth, source sum, Euler/clamp, coalesced stores + periodic frame).
//   * the 256-point transform is 16 x 16 in registers: 16 lanes per row, 16 samples per lane, ONE shared-memory exchange
//     (padded rows: every access is base + immediate), constants of the 16-point kernels as immediates; two rows per warp,
//     __syncwarp only.
// HBM traffic per cell and step: the state once (+14 % strip halo and the band halo, mostly L2 hits), the Euler operand (an L2
// hit: the same CTA loaded that row <= 38 rows earlier) and the output: 1.0x the algorithmic 8 B at 16384^2 (ncu).
// tools/march_proto.py is the numpy model of the index arithmetic below.
// Code size: everything below is inlined into the kernel (one copy of the forward transform: the prologue runs as extra rounds of
// the marching loop).  A version with out-of-line transforms / frame routines (3.8 K instructions) measured 20-50 % slower: the
// calls cost registers and spills around every call site.  Not kept.
#include "common.cuh"

// A/B switches of the developer builds (tools/march_ab.sh); the defaults are what ships
#ifndef MARCH_PACKED_MUL
#define MARCH_PACKED_MUL 0   // complex x constant as FMUL2 + FFMA2 on a swapped pair (0: four scalar instructions)
#endif
#ifndef MARCH_PACKED_ROT
#define MARCH_PACKED_ROT 0   // the +-i legs of the radix-4 butterfly as FFMA2 on a swapped pair (0: four scalar adds)
#endif
#ifndef MARCH_PREFETCH
#define MARCH_PREFETCH 0     // forward-row loads issued before the FIR, Euler operands before the transforms
#endif
#ifndef MARCH_PREFETCH_EPI
#define MARCH_PREFETCH_EPI 1 // Euler operands issued before the transforms (16 registers held across them)
#endif
#ifndef MARCH_ONE_FORWARD
#define MARCH_ONE_FORWARD 1  // the prologue runs as extra rounds of the marching loop: ONE inlined copy of the forward transform
#endif
#ifndef MARCH_MAXLP
#define MARCH_MAXLP 128      // rows pairs per band at most (taller bands: fewer redundant forward transforms, but too few CTAs)
#endif
#ifndef MARCH_FORCELP
#define MARCH_FORCELP 0      // != 0: fixed band height (A/B runs)
#endif
#ifndef MARCH_RELOAD_TAPS
#define MARCH_RELOAD_TAPS 1  // 1: the single-destination kernel reloads its 16 tap pairs before every FIR (32 registers free elsewhere)
#endif
#ifndef MARCH_MULTI_XEDGE
#define MARCH_MULTI_XEDGE 1
#endif
#ifndef MARCH_PRO_UNITS
#define MARCH_PRO_UNITS 4
#endif
#ifndef MARCH_EARLY_U
#define MARCH_EARLY_U 0      // the epilogue takes its 8 filtered samples into registers first; the barrier that frees `filt` for the
#endif                       // next FIR follows those loads instead of the whole epilogue (its global loads / stores overlap the FIR)
#ifndef MARCH_L2_PREFETCH
#define MARCH_L2_PREFETCH 1  // the forward warps ask L2 for the rows they will transform after the FIR (prefetch.global.L2: no registers held)
#endif
#ifndef MARCH_EDGE_UNITS
#define MARCH_EDGE_UNITS 2
#endif
#ifndef MARCH_SKEW
#define MARCH_SKEW 0         // per-mille share of a co-resident CTA pair's rows given to the first-placed CTA (see next_band; 0: even split)
#endif
#ifndef MARCH_JOB_WAVES
#define MARCH_JOB_WAVES 1    // balanced launches: grid = this many waves of CTAs (0: fixed-height bands for every launch)
#endif

namespace b200ca {

namespace march {

constexpr int F = B200CA_FRAME;
constexpr int R = 15;
constexpr int KROWS = R + 1;
constexpr int N = 256;            // transform length
constexpr int HALO = 16;          // samples discarded on each side of a strip (>= R)
constexpr int V = N - 2 * HALO;   // valid outputs per strip
constexpr int FT = 8;             // row pairs per iteration
constexpr int NBLK = 5;           // ring blocks of FT rows: 40 rows >= FT + 2R
constexpr int RING = NBLK * FT;
constexpr int ROW4 = N / 2;       // float4 of a spectra row (tap-table stride)
// Shared-memory rows are PADDED so that every access is base + immediate offset (an XOR swizzle costs two integer
// instructions per access -- 27 % of the kernel's instructions in the first version):
//   slot layout      float4 index 9*ka + j        (lane ka's 8 frequency pairs; quarter-warp conflict free)
//   exchange layout  float2 index 17*t + r        (16 x 16 transpose between the two 16-point stages)
//   sample layout    float2 index n               (inverse output: (row a, row b) of sample n)
constexpr int ROWP4 = 144;        // float4 per padded row (2304 B >= 16*17 float2)
constexpr int THREADS = 256;
constexpr int EDGE_UNITS = MARCH_EDGE_UNITS;    // split launches: what the slow rounds at a column's first / last rows cost their band, in rounds
constexpr int PRO_UNITS = MARCH_PRO_UNITS;      // what entering a column costs a run, in rounds of the marching loop: the prologue + the slow rounds at a world's first / last rows (fitted: 0..8 tried)
constexpr size_t SMEM = (size_t)(RING + FT) * ROWP4 * sizeof(float4) + 16 * 17 * sizeof(float2);

struct Args {
    const float *state_in;
    float *out;
    const float2 *kf;   // (B, NS, C, KROWS, 128) tap pairs in FIR-thread order (march_prepare_kernel)
    const float *mu, *sigma, *weights;
    int B, C, NS, k_mult, H, W, pitch;
    int Lp, niter;      // banded launches: row pairs per band (multiple of FT), Lp / FT
    int band0;          // banded launches: first band of this launch (a caller that overlaps an exchange issues its boundary bands first)
    int jobs;           // 1: balanced launch -- the grid is ONE wave of CTAs that share the (world, strip, 16-row unit) space evenly
    int units, strips;  // 16-row units per column of bands (ceil(H / 16)), strips per world
    int split_k;        // balanced launches: != 0: every column is cut into split_k equal bands, one CTA each (no run crosses a column)
    int skew;           // balanced launches on a full wave: per-mille share of a CTA pair's rows that goes to the first-placed CTA (0: even)
    int src;            // source-kernel index i*k_mult + m of this launch
    int acc_in;         // 1: add to the partial sums already in `out` (sources after the first)
    int final_src;      // 1: last source -> Euler step / clamp (LENIA) or plain sum (AFF)
    float dt;
    int mode, row_wrap;
    // Row slab of a torus sharded over GPUs (NVLink P2P halo, no host-side exchange; all null / 0 otherwise).  The slab's first
    // and last F rows are ALSO stored into the neighbours' ghost rows (their `out` buffer, mapped into this address space),
    // then this CTA bumps the neighbour's arrival counter; CTAs of the first / last band wait until the counters of THIS rank
    // say that the ghost rows they are about to read have arrived.  The CTAs that read a rank's ghost rows are the same CTAs
    // that produce the rows sent back to the writer of those ghost rows, so the one counter orders both directions
    // (read-after-write of the halo, write-after-read of the buffer that is reused two steps later).
    float *peer_up_out, *peer_dn_out;            // out buffers of the ranks owning the rows above / below (framed, same shape)
    unsigned *sig_up, *sig_dn;                   // their counters: "your bottom ghost rows arrived" / "your top ghost rows arrived"
    const unsigned *wait_top, *wait_bot;         // this rank's counters for its top / bottom ghost rows
    unsigned wait_target;                        // arrivals expected before this step may read its ghost rows
    int nbands;                                  // bands of the whole slab (blockIdx.y is permuted: boundary bands first)
    unsigned *err_flag;                          // set if a wait timed out (no hang: the step then reads stale ghost rows)
#ifdef B200CA_DEV
    unsigned long long *trace;                   // developer builds: per CTA (start ns, end ns, SM id) of the last launch, or null
#endif
};

// Every fp32 instruction with three register operands -- scalar or packed -- occupies the FMA pipe for two cycles per warp
// (ncu: pipe time = 2 x the fp instruction count), so the transforms are written in PACKED fp32x2 form throughout: a complex
// number is one register pair, an add / real-scale / multiply-accumulate is one FADD2 / FMUL2 / FFMA2.  Products with i
// need the swapped pair (a.y, a.x): two MOVs on the ALU pipe, which has room.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
// a - b as ONE packed instruction (FADD2 has no negate modifier)
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ float2 cswap(float2 a) { return make_float2(a.y, a.x); }
// a * (c + i s) = a * (c, c) + (a.y, a.x) * (-s, s)
__device__ __forceinline__ float2 cmul_cs(float2 a, float c, float s) {
#if MARCH_PACKED_MUL
    return __ffma2_rn(cswap(a), make_float2(-s, s), __fmul2_rn(a, make_float2(c, c)));
#else
    return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
#endif
}

// 4-point DFT, natural order in and out; INV: e^{+i...}
template <bool INV>
__device__ __forceinline__ void dft4(float2 &a, float2 &b, float2 &c, float2 &d) {
    // t1 +- (b-d) * (-+i) = t1 +- (bd.y, -bd.x)  (forward)
    constexpr float S = INV ? -1.f : 1.f;
#if MARCH_PACKED_ROT
    const float2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), bd = cswap(csub(b, d));
    a = cadd(t0, t2);
    c = csub(t0, t2);
    b = __ffma2_rn(bd, make_float2(S, -S), t1);
    d = __ffma2_rn(bd, make_float2(-S, S), t1);
#else
    const float2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), bd = csub(b, d);
    a = cadd(t0, t2);
    c = csub(t0, t2);
    b = make_float2(t1.x + S * bd.y, t1.y - S * bd.x);
    d = make_float2(t1.x - S * bd.y, t1.y + S * bd.x);
#endif
}

// 16-point DFT on registers, natural order in and out (two radix-4 stages, constants as immediates).
// V[ka + 4 kb] = sum_{q1} (-+i)^{q1 kb} w^{q1 ka} sum_{q2} v[q1 + 4 q2] (-+i)^{q2 ka},  w = exp(-+2 pi i / 16)
template <bool INV>
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, C2 = 0.70710678118654752f;
    constexpr float SG = INV ? 1.f : -1.f;
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1) dft4<INV>(v[q1], v[q1 + 4], v[q1 + 8], v[q1 + 12]);  // v[q1 + 4 ka]
    // twiddles w^{q1 ka}: exponents 1,2,3 / 2,4,6 / 3,6,9
    v[1 + 4] = cmul_cs(v[1 + 4], C1, SG * S1);
    v[1 + 8] = cmul_cs(v[1 + 8], C2, SG * C2);
    v[1 + 12] = cmul_cs(v[1 + 12], S1, SG * C1);
    v[2 + 4] = cmul_cs(v[2 + 4], C2, SG * C2);
#if MARCH_PACKED_MUL
    v[2 + 8] = __fmul2_rn(cswap(v[2 + 8]), make_float2(-SG, SG));  // * (-+i)
#else
    v[2 + 8] = INV ? make_float2(-v[2 + 8].y, v[2 + 8].x) : make_float2(v[2 + 8].y, -v[2 + 8].x);
#endif
    v[2 + 12] = cmul_cs(v[2 + 12], -C2, SG * C2);
    v[3 + 4] = cmul_cs(v[3 + 4], S1, SG * C1);
    v[3 + 8] = cmul_cs(v[3 + 8], -C2, SG * C2);
    v[3 + 12] = cmul_cs(v[3 + 12], -C1, -SG * S1);
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) dft4<INV>(v[4 * ka], v[4 * ka + 1], v[4 * ka + 2], v[4 * ka + 3]);  // V[ka + 4 kb] at v[4 ka + kb]
    // transpose the 4 x 4 register grid back to natural order (pure renaming once unrolled)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) {
            const float2 tmp = v[4 * a + b];
            v[4 * a + b] = v[4 * b + a];
            v[4 * b + a] = tmp;
        }
}

// One forward transform by a 16-lane group (lane t): image rows `ra` (re) and `rb` (im), 256 samples from framed column
// `col0` on; the spectrum goes to `row4` in slot layout: float4 index 9*ka + j = (re, re', im, im') of the
// frequencies ka + 16*(2j), ka + 16*(2j+1).  The 16 x 16 exchange runs in place in the destination row.
template <bool COHERENT>
__device__ __forceinline__ float ld_state(const float *p) {
#ifdef MARCH_SLAB_LDG
    return __ldg(p);
#else
    return COHERENT ? __ldcg(p) : __ldg(p);
#endif   // ghost rows written by a peer GPU during this kernel: L2 path, not the read-only one
}

template <bool COHERENT>
__device__ __forceinline__ void forward_load(const float *__restrict__ ra, const float *__restrict__ rb, int col0, int pitch, int t,
                                             float2 (&v)[16]) {
    ra += col0 + t;
    rb += col0 + t;
    if (col0 + N <= pitch) {  // every strip but the last: plain loads at immediate offsets
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = make_float2(ld_state<COHERENT>(ra + 16 * q), ld_state<COHERENT>(rb + 16 * q));
    } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const bool ok = col0 + t + 16 * q < pitch;
            v[q] = make_float2(ok ? ld_state<COHERENT>(ra + 16 * q) : 0.f, ok ? ld_state<COHERENT>(rb + 16 * q) : 0.f);
        }
    }
}

__device__ __forceinline__ void forward_compute(float2 (&v)[16], float4 *row4, const float2 *tw, int t) {
    dft16<false>(v);  // v[ka] = sum_q x[t + 16 q] W16^{q ka}
    float2 *row2 = reinterpret_cast<float2 *>(row4);
#pragma unroll
    for (int ka = 1; ka < 16; ++ka) {
        const float2 w = tw[17 * t + ka];  // W256^{t ka}
        v[ka] = cmul_cs(v[ka], w.x, w.y);
    }
#pragma unroll
    for (int ka = 0; ka < 16; ++ka) row2[17 * t + ka] = v[ka];
    __syncwarp();
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) v[n1] = row2[17 * n1 + t];  // lane t now owns ka = t
    __syncwarp();
    dft16<false>(v);  // v[kb] = X[t + 16 kb]
#pragma unroll
    for (int j = 0; j < 8; ++j) row4[9 * t + j] = make_float4(v[2 * j].x, v[2 * j + 1].x, v[2 * j].y, v[2 * j + 1].y);
}

// One inverse transform by a 16-lane group, in place: filtered spectrum (slot layout) -> float2 slot n = (out row a, out row b)
__device__ __forceinline__ void inverse_row(float4 *row4, const float2 *tw, int t) {
    float2 v[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 z = row4[9 * t + j];
        v[2 * j] = make_float2(z.x, z.z);
        v[2 * j + 1] = make_float2(z.y, z.w);
    }
    __syncwarp();    // every lane has its spectrum before the row is reused for the exchange
    dft16<true>(v);  // v[n1] = sum_kb Y[t + 16 kb] W16^{-n1 kb}
    float2 *row2 = reinterpret_cast<float2 *>(row4);
#pragma unroll
    for (int n1 = 1; n1 < 16; ++n1) {
        const float2 w = tw[17 * t + n1];
        v[n1] = cmul_cs(v[n1], w.x, -w.y);  // conj(W256^{t n1})
    }
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) row2[17 * t + n1] = v[n1];
    __syncwarp();
#pragma unroll
    for (int ka = 0; ka < 16; ++ka) v[ka] = row2[17 * ka + t];  // lane t now owns n1 = t
    __syncwarp();
    dft16<true>(v);  // v[n2] = y[t + 16 n2]
#pragma unroll
    for (int q = 0; q < 16; ++q) row2[t + 16 * q] = v[q];
}

// Column FIR of 4 output row pairs (4*H4 .. 4*H4+3 of the current block of 8) for the thread's two frequencies.  The 34 input
// rows start at row 4*H4 of ring block b0 and run through blocks b0 .. b0+4; `blk[i]` points at this thread's float4 of the
// first row of block (b0 + i) % 5, so every load is pointer + immediate offset (no per-row wrap arithmetic).
template <int H4>
__device__ __forceinline__ void fir_half(const float4 *const (&blk)[NBLK], const float2 (&kt)[KROWS], float4 *fout) {
    float2 are[4], aim[4];
#pragma unroll
    for (int y = 0; y < 4; ++y) are[y] = aim[y] = make_float2(0.f, 0.f);
#pragma unroll
    for (int rr = 0; rr < 4 + 2 * R; ++rr) {
        constexpr int unused = 0;
        (void)unused;
        const int row = 4 * H4 + rr;
        const float4 z = blk[row / FT][(row % FT) * ROWP4];
        const float2 zre = make_float2(z.x, z.y), zim = make_float2(z.z, z.w);
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int dy = rr - R - y;
            if (dy >= -R && dy <= R) {
                are[y] = __ffma2_rn(kt[dy < 0 ? -dy : dy], zre, are[y]);
                aim[y] = __ffma2_rn(kt[dy < 0 ? -dy : dy], zim, aim[y]);
            }
        }
    }
#pragma unroll
    for (int y = 0; y < 4; ++y) fout[y * ROWP4] = make_float4(are[y].x, are[y].y, aim[y].x, aim[y].y);
}

struct EpiArgs {
    const float2 *u;       // this thread's sample of output pair 0 (stride 2*ROWP4 float2 per pair)
    // Row addresses are byte pointers + r * pitchB with a 32-bit product: ONE IMAD.WIDE per address (element-indexed 64-bit
    // arithmetic cost four integer instructions per row: 128 of the ~1000 instructions of a marching round).
    const char *sa, *sb;   // state, first row of the first / second half of this block of 8 row pairs (EULER)
    char *oa, *ob;         // output, same rows
    int pitchB;            // row pitch in bytes
    int pitch, H, W, ya0, Lp, xw, row_wrap;
    int xoff;              // distance to this column's periodic x image (+-W for the 16 columns next to the world edge, else 0)
    float mu, nc2, w2, wn, dt;   // growth(u) * weight = exp2((u - mu)^2 * nc2) * w2 + wn
    float *peer_up, *peer_dn;    // destination plane in the neighbours' out buffers (row slabs), or null
};

__device__ __forceinline__ const float *row_ptr(const char *base, int r, int pitchB) { return reinterpret_cast<const float *>(base + (long long)(r * pitchB)); }
__device__ __forceinline__ float *row_ptr(char *base, int r, int pitchB) { return reinterpret_cast<float *>(base + (long long)(r * pitchB)); }

// a row slab's first / last F rows go to the neighbour's bottom / top ghost rows as well (with their periodic x images)
__device__ __forceinline__ void store_peer(const EpiArgs &e, int y, float v) {
    float *dst = nullptr;
    if (e.peer_up && y < F) dst = e.peer_up + (size_t)(e.H + F + y) * e.pitch;
    else if (e.peer_dn && y >= e.H - F) dst = e.peer_dn + (size_t)(y - e.H + F) * e.pitch;
    if (!dst) return;
    dst[F + e.xw] = v;
    if (e.xw < F) dst[F + e.xw + e.W] = v;
    if (e.xw >= e.W - F) dst[F + e.xw - e.W] = v;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// the Euler operands (state rows a and b of 8 row pairs), issued long before they are needed
template <int EDGE>
__device__ __forceinline__ void epilogue_prefetch(const EpiArgs &e, float2 (&st)[FT]) {
#pragma unroll
    for (int r = 0; r < FT; ++r) {
        const bool oka = EDGE < 2 || e.ya0 + r < e.H, okb = EDGE < 2 || e.ya0 + e.Lp + r < e.H;
        st[r] = make_float2(oka ? __ldg(row_ptr(e.sa, r, e.pitchB)) : 0.f, okb ? __ldg(row_ptr(e.sb, r, e.pitchB)) : 0.f);
    }
}

// growth + weighted source sum (+ partial sums of earlier sources) (+ Euler step / clamp) and the stores of 8 row pairs;
// the two rows of a pair are the two halves of packed fp32x2 operations.
// EDGE = 0: a block of 8 row pairs whose cells have no periodic images and whose rows all lie inside the world;  1: the same rows in
// the first / last strip of a world (2 of the 5 strips of a 1024-wide world): the cells within 16 columns of the world edge store
// their periodic x image as well (one predicated store per row -- as cheap as EDGE = 0, which matters because a balanced launch
// ends when its slowest CTA does);  2: the general path (first / last 16 rows of a torus or a slab, rows beyond the world).
template <int EDGE, bool ACC, bool EULER, bool EARLY>
__device__ __forceinline__ void epilogue(const EpiArgs &e, const float2 (&st)[FT], const float2 (&uu)[FT]) {
    float2 pv[FT];
    if (ACC) {
#pragma unroll
        for (int r = 0; r < FT; ++r) {
            const bool oka = EDGE < 2 || e.ya0 + r < e.H, okb = EDGE < 2 || e.ya0 + e.Lp + r < e.H;
            pv[r] = make_float2(oka ? *row_ptr(e.oa, r, e.pitchB) : 0.f, okb ? *row_ptr(e.ob, r, e.pitchB) : 0.f);
        }
    }
    const float2 nmu = make_float2(-e.mu, -e.mu), nc2 = make_float2(e.nc2, e.nc2), w2 = make_float2(e.w2, e.w2), wn = make_float2(e.wn, e.wn),
                 dt2 = make_float2(e.dt, e.dt);
#pragma unroll
    for (int r = 0; r < FT; ++r) {
        const float2 d = __fadd2_rn(EARLY ? uu[r] : e.u[r * 2 * ROWP4], nmu);
        const float2 q = __fmul2_rn(__fmul2_rn(d, d), nc2);
        float2 v = __ffma2_rn(make_float2(ex2_approx(q.x), ex2_approx(q.y)), w2, wn);
        if (ACC) v = __fadd2_rn(v, pv[r]);
        if (EULER) {
            v = __ffma2_rn(dt2, v, st[r]);
            v = make_float2(fminf(fmaxf(v.x, 0.f), 1.f), fminf(fmaxf(v.y, 0.f), 1.f));
        }
        if (EDGE == 0) {
            *row_ptr(e.oa, r, e.pitchB) = v.x;
            *row_ptr(e.ob, r, e.pitchB) = v.y;
        } else if (EDGE == 1) {
            float *pa = row_ptr(e.oa, r, e.pitchB), *pb = row_ptr(e.ob, r, e.pitchB);
            *pa = v.x;
            *pb = v.y;
            if (e.xoff) {
                pa[e.xoff] = v.x;
                pb[e.xoff] = v.y;
            }
        } else {
            // rows beyond the world are skipped; the first / last 16 rows of a torus store their periodic y image (and the corner
            // image) as well, the first / last 16 rows of a slab go to the neighbour's ghost rows.  All inline: a balanced launch
            // ends with its slowest CTA, and the out-of-line image routine made the four rounds per column that touch these rows
            // cost as much as eight.
            const int ya = e.ya0 + r, yb = ya + e.Lp;
            const int xoff = e.xoff;
            if (ya < e.H) {
                float *p = row_ptr(e.oa, r, e.pitchB);
                *p = v.x;
                if (xoff) p[xoff] = v.x;
                const int yoff = !e.row_wrap ? 0 : (ya < F ? e.H : (ya >= e.H - F ? -e.H : 0));
                if (yoff) {
                    float *q = p + (long long)yoff * e.pitch;
                    *q = v.x;
                    if (xoff) q[xoff] = v.x;
                }
                store_peer(e, ya, v.x);
            }
            if (yb < e.H) {
                float *p = row_ptr(e.ob, r, e.pitchB);
                *p = v.y;
                if (xoff) p[xoff] = v.y;
                const int yoff = !e.row_wrap ? 0 : (yb < F ? e.H : (yb >= e.H - F ? -e.H : 0));
                if (yoff) {
                    float *q = p + (long long)yoff * e.pitch;
                    *q = v.y;
                    if (xoff) q[xoff] = v.y;
                }
                store_peer(e, yb, v.y);
            }
        }
    }
}

// The next band of CTA `cta` in a balanced launch (shared by the kernel and the host-side plan, b200ca_lenia_march_plan_bands):
// `done` counts what the CTA has consumed so far (virtual units of its run, or 1 once its single split band is out).
// split_k != 0: every column of bands is cut into split_k equal bands (chosen by the host when that leaves no CTA with more rows
// than a run would get: e.g. 74 columns on 296 CTAs).  split_k == 0: runs that are equal in COST -- a column is `units` rounds of
// the marching loop plus PRO_UNITS rounds' worth of prologue, so a run that crosses into the next column (a second prologue)
// gets fewer rows than one that does not.
// `skew` (per mille, 0 = off, and off by default): on a full wave the two CTAs that share an SM do not share it evenly.  The per-CTA
// trace (tools/march_trace.py) shows, on every one of the 148 SMs, one CTA running at its SOLO rate and its sibling on the issue
// slots that are left (0.58 : 0.42 of the SM), so that the favoured CTA ends at 916 us of a 16384^2 step and the other one
// finishes alone at 1173 us, well under the SM's two-CTA throughput: 9 % of the launch.  With `skew` the grid is handled as
// grid/2 PAIRS -- CTA i and CTA i + grid/2, the block scheduler deals blockIdx 0 .. nSM-1 one per SM first -- and a pair's share
// of the rows is cut skew : 1000 - skew.  Measured: 540 gains 2 % on the batch and 0.4 % at 16384^2; at 580, where both CTAs
// should end together, which CTA is favoured flips from SM to SM (it follows the warp slots freed by the PREVIOUS launch) and
// the step is 15 % slower.  Left off; the robust fix is a pair of sub-CTAs inside one 512-thread CTA that split a band's
// remaining rounds when the first of them runs out of work.
__host__ __device__ __forceinline__ bool next_band(int units, long long ncols, int grid, int split_k, int skew, int cta, int &done, int &col,
                                                   int &u0, int &nu) {
    const int half = grid >> 1;
    const bool paired = skew > 0;
    const int pair = paired ? (cta < half ? cta : cta - half) : 0;
    const bool fast = cta < half;
    if (split_k) {
        // (the bands that hold a column's first / last rows run EDGE_UNITS rounds' worth of slow edge rounds -- periodic y images,
        // a slab's halo stores and waits -- and get that many units less)
        if (done) return false;
        done = 1;
        const int e = split_k > 1 && units >= 8 * split_k ? EDGE_UNITS : 0;
        const long long vcol = units + 2 * e;
        long long v0, v1;   // this CTA's piece of the column's virtual units [0, vcol)
        if (paired) {       // split_k even: split_k / 2 pairs per column
            const int ppc = split_k >> 1;
            col = pair / ppc;
            const int pp = pair - col * ppc;
            const long long lo = vcol * pp / ppc, hi = vcol * (pp + 1) / ppc, mid = lo + (hi - lo) * skew / 1000;
            v0 = fast ? lo : mid;
            v1 = fast ? mid : hi;
        } else {
            col = cta / split_k;
            const int part = cta - col * split_k;
            v0 = vcol * part / split_k;
            v1 = vcol * (part + 1) / split_k;
        }
        const int b0 = v0 <= 0 ? 0 : (int)v0 - e, b1 = v1 >= vcol ? units : (int)v1 - e;
        u0 = b0 < 0 ? 0 : b0;
        nu = (b1 > units ? units : b1) - u0;
        return nu > 0;
    }
    const int vunits = units + PRO_UNITS;
    const long long total = (long long)vunits * ncols;
    long long v_begin, v_end;
    if (paired) {
        const long long lo = total * pair / half, hi = total * (pair + 1) / half, mid = lo + (hi - lo) * skew / 1000;
        v_begin = fast ? lo : mid;
        v_end = fast ? mid : hi;
    } else {
        v_begin = total * cta / grid;
        v_end = total * (cta + 1) / grid;
    }
    for (;;) {
        const long long v = v_begin + done;
        if (v >= v_end) return false;
        col = (int)(v / vunits);
        const int v0 = (int)(v - (long long)col * vunits);
        const long long left = v_end - v;
        const int nv = (long long)(vunits - v0) < left ? vunits - v0 : (int)left;
        done += nv;
        u0 = v0 > PRO_UNITS ? v0 - PRO_UNITS : 0;
        nu = (v0 + nv > PRO_UNITS ? v0 + nv - PRO_UNITS : 0) - u0;
        if (nu > 0) return true;
    }
}

// block 256.  SINGLE: one destination plane (C == 1): taps stay in registers.
// Two ways to hand out the bands:
//   * balanced (a.jobs, every whole-world / whole-slab step): grid = ONE wave of CTAs (2 per SM).  The columns of bands -- (world,
//     strip) pairs, each ceil(H / 16) units of 16 rows tall -- are laid end to end and cut into gridDim.x equal runs of units; a CTA
//     marches down its run, starting a new band (prologue: 32 extra forward transforms) only at its first unit and where the run
//     crosses into the next column.  Fixed-height bands had to choose between tall bands that leave the last wave of CTAs mostly
//     empty and short ones that double the forward transforms (24 x 1024^2: 1920 CTAs of 64 rows = 6.5 waves, 2x the forward
//     work); now every SM is busy for the same time with ~1.1x.
//   * banded (band_begin / band_end of b200ca_lenia_step_march): grid (strips, bands of this launch, B), fixed band height.
template <bool SINGLE>
__global__ void __launch_bounds__(THREADS, 2) lenia_march_kernel(const Args a) {
    extern __shared__ __align__(16) float4 msm[];
    float4 *ring = msm;                                  // RING rows of ROWP4 float4
    float4 *filt = msm + RING * ROWP4;                   // FT rows: filtered spectra -> inverse exchange -> outputs
    float2 *tw = reinterpret_cast<float2 *>(filt + FT * ROWP4);  // W256^{t r} at [17 t + r]
    const int tid = threadIdx.x, lane16 = tid & 15, grp = tid >> 4;  // 16 groups of 16 lanes
    const size_t plane_elems = (size_t)(a.H + 2 * F) * a.pitch;
    const int max_frow = a.H + 2 * F - 1;
    const bool slab = a.wait_top != nullptr || a.wait_bot != nullptr;

    // input-independent prologue: twiddle table
    {
        const int t = tid >> 4, r = tid & 15;
        float sn, cs;
        sincospif(-2.f * (float)((t * r) & 255) / 256.f, &sn, &cs);
        tw[17 * t + r] = make_float2(cs, sn);
    }
    const int f = tid & 127, h = tid >> 7;
    const int fslot = 9 * (f >> 3) + (f & 7);            // this FIR thread's float4 in a padded spectra row
    // single destination only: with the partial sums of earlier sources in registers as well the multi-destination kernel spills
    constexpr bool EARLY_U = MARCH_EARLY_U && SINGLE;
    const bool lenia = a.mode == B200CA_MODE_LENIA;
    const bool edge_rows = a.row_wrap || a.peer_up_out || a.peer_dn_out;
    pdl_wait_then_release();  // the state is the previous step's output
    __syncthreads();          // twiddle table
#ifdef B200CA_DEV
    if (a.trace && tid == 0) {
        unsigned long long t0;
        unsigned sm;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        a.trace[3 * blockIdx.x] = t0;
        a.trace[3 * blockIdx.x + 2] = sm;
    }
#endif

    int done = 0;             // balanced launches: units of this CTA's run that are finished
#pragma unroll 1
    for (;;) {
        // ---- the next band of this CTA
        int strip, b, y0, Lp, niter;
        if (a.jobs) {
            int col, u0;
            if (!next_band(a.units, (long long)a.strips * a.B, (int)gridDim.x, a.split_k, a.skew, (int)blockIdx.x, done, col, u0, niter)) break;
            strip = col % a.strips;
            b = col / a.strips;
            y0 = 2 * FT * u0;
            Lp = FT * niter;
        } else {
            if (done) break;
            done = 1;
            strip = blockIdx.x;
            b = blockIdx.z;
            // callers that overlap a halo exchange: boundary bands first (blockIdx.y = 0 -> first band, 1 -> last band)
            int band = a.band0 + blockIdx.y;
#ifndef MARCH_NOPERM
            if (a.nbands > 1 && (a.sig_up || a.sig_dn)) band = blockIdx.y == 0 ? 0 : (blockIdx.y == 1 ? a.nbands - 1 : (int)blockIdx.y - 1);
#endif
            Lp = a.Lp;
            niter = a.niter;
            y0 = band * 2 * Lp;
        }
        const bool top = y0 == 0, bot = y0 + 2 * Lp >= a.H;   // this band holds the first / last rows of the world (slab)
        const int col0 = strip * V;                          // framed column of sample n = 0 (world column strip*V - HALO)
        const int src_plane = b * a.C + a.src / a.k_mult;
        const float *splane = a.state_in + (size_t)src_plane * plane_elems;

        // taps of the first destination
        float2 kt[KROWS];
        const float2 *kf0 = a.kf + (size_t)((b * a.NS + a.src) * a.C) * KROWS * ROW4 + f;
        if (SINGLE && !MARCH_RELOAD_TAPS) {
#pragma unroll
            for (int dy = 0; dy < KROWS; ++dy) kt[dy] = __ldg(kf0 + dy * ROW4);
        }
        if (slab && (top || bot)) {
            if (tid == 0) {
                // ghost rows written by the neighbours' previous step: wait for their arrival counters (bounded: a lost neighbour must
                // not hang the GPU -- after ~4 s the error flag is raised and the step goes on)
                unsigned long long t0;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                for (int side = 0; side < 2; ++side) {
                    const unsigned *cnt = side == 0 ? a.wait_top : a.wait_bot;
                    const bool mine = side == 0 ? top : bot;
                    if (!cnt || !mine) continue;
                    for (;;) {
                        unsigned v;
                        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
                        if ((int)(v - a.wait_target) >= 0) break;
                        unsigned long long t1;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                        if (t1 - t0 > 4000000000ull) {
                            if (a.err_flag) atomicExch(a.err_flag, 1u);
                            break;
                        }
                        __nanosleep(200);
                    }
                }
            }
            __syncthreads();      // ghost rows
        }

        // spectra row z of the band <- image rows (y0 + z - R, y0 + Lp + z - R), clamped into the framed plane
        auto forward_fetch = [&](int z, float2 (&v)[16]) {
            const int fa = min(y0 + z - R + F, max_frow), fb = min(y0 + Lp + z - R + F, max_frow);
            // ghost rows (the frame rows of a slab) are written by a peer GPU while this kernel may already run: they are read
            // through L2 (ld.global.cg); everything else takes the read-only path, whose L1 merges the 64-byte row pieces
            // (all loads through L2 cost 12 % of the step)
            if (slab && (fa < F || fa >= a.H + F || fb >= a.H + F)) forward_load<true>(splane + (size_t)fa * a.pitch, splane + (size_t)fb * a.pitch, col0, a.pitch, lane16, v);
            else forward_load<false>(splane + (size_t)fa * a.pitch, splane + (size_t)fb * a.pitch, col0, a.pitch, lane16, v);
        };
#if !MARCH_ONE_FORWARD
        // prologue: blocks 0..4 (40 rows) by all 16 groups
#pragma unroll 1
        for (int z = grp; z < RING; z += 16) {
            float2 v[16];
            forward_fetch(z, v);
            forward_compute(v, ring + (z % RING) * ROWP4, tw, lane16);
        }
#endif

        // growth(u) * weight = exp2((u - mu)^2 * nc2) * w2 + wn  (lenia.py:423-426: 2 exp(-(u-mu)^2 / (2 sigma^2)) - 1); one destination:
        // constant for the whole band
        float g_mu = 0.f, g_nc2 = 0.f, g_w2 = 0.f, g_wn = 0.f;
        if (SINGLE) {
            const int pidx = (b * a.NS + a.src) * a.C;
            const float sg = __ldg(a.sigma + pidx), wt = __ldg(a.weights + pidx);
            g_mu = __ldg(a.mu + pidx);
            g_nc2 = -1.4426950408889634f / (2.f * sg * sg);
            g_w2 = 2.f * wt;
            g_wn = -wt;
        }
        const int xw = col0 - HALO + tid;                    // world column of this thread's sample n = tid (epilogue)
        const bool col_ok = tid >= HALO && tid < N - HALO && xw < a.W;
        // CTAs that own cells with periodic images (first / last columns, first / last rows of a torus) or a partial last band
        // (decided per block of 8 row pairs: in a 256-row band only the first / last 16 rows of the world take the slow path)
        const bool edge_x = strip == 0 || col0 - HALO + N - HALO > a.W - F;
        // MARCH_ONE_FORWARD: rounds it = -3, -2, -1 are the prologue -- all 16 groups run forward transforms (blocks 0..4 of the ring,
        // 40 rows) through the SAME inlined code the marching rounds use (a second copy in a separate prologue loop costs 18 KB of
        // instruction-cache footprint, which short CTAs pay for: ncu `no_instruction` 25 % of the stall samples on the 24 x 1024^2 batch)
        constexpr int PRO = MARCH_ONE_FORWARD ? (RING + 15) / 16 : 0;
#pragma unroll 1
        for (int it = -PRO; it < niter; ++it) {
            const bool pro = it < 0;
            const int ra0 = y0 + FT * it, rb0 = ra0 + Lp;   // first rows of the two halves of this block
            const bool edge_y = rb0 + FT > a.H || (edge_rows && (ra0 < F || rb0 + FT > a.H - F || (ra0 + FT > a.H - F) || rb0 < F));
            const int edge = edge_y ? 2 : (edge_x ? (SINGLE || MARCH_MULTI_XEDGE ? 1 : 2) : 0);
#pragma unroll 1
            for (int j = (SINGLE || !pro) ? 0 : a.C - 1; j < (SINGLE ? 1 : a.C); ++j) {
                const int pidx = (b * a.NS + a.src) * a.C + j;
                if ((!SINGLE || MARCH_RELOAD_TAPS) && !pro) {
                    const float2 *kfj = kf0 + (size_t)j * KROWS * ROW4;
#pragma unroll
                    for (int dy = 0; dy < KROWS; ++dy) kt[dy] = __ldg(kfj + dy * ROW4);
                }
                // the forward transform this group runs after the FIR (groups 8-15, last destination): its 32 loads are issued
                // now and stay in flight behind the FIR
                const int zf = pro ? grp + 16 * (it + PRO) : FT * (it + NBLK) + (grp - FT);
                const bool do_fwd = pro ? zf < RING : (grp >= FT && j == (SINGLE ? 0 : a.C - 1) && it + 1 < niter);
                float2 vf[16];
#if MARCH_PREFETCH
                if (do_fwd) forward_fetch(zf, vf);
#endif
#if MARCH_L2_PREFETCH
                if (do_fwd && !pro) {
                    // 2 rows x 1 KB = 16 lines of 128 B: one per lane of the group (row a: lanes 0-7, row b: lanes 8-15)
                    const int frow = min(y0 + ((lane16 & 8) ? Lp : 0) + zf - R + F, max_frow);
                    const int colp = min(col0 + 32 * (lane16 & 7), a.pitch - 1);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(splane + (size_t)frow * a.pitch + colp));
                }
#endif
                // ring rows of this iteration are complete; the previous epilogue is done with `filt` (MARCH_EARLY_U: it said so itself,
                // right after loading its samples -- only a band's first FIR still has to wait for the prologue's ring rows)
                if (!pro && (!EARLY_U || (it == 0 && j == 0))) __syncthreads();
                // ---- FIR: output pairs 4h .. 4h+3 of this block from ring rows (8 it + 4h) .. (+33), two frequencies per thread
                if (!pro) {
                    const float4 *blk[NBLK];
                    int bb = it % NBLK;
#pragma unroll
                    for (int i = 0; i < NBLK; ++i) {
                        blk[i] = ring + bb * FT * ROWP4 + fslot;
                        bb = bb + 1 == NBLK ? 0 : bb + 1;
                    }
                    if (h == 0) fir_half<0>(blk, kt, filt + fslot);
                    else fir_half<1>(blk, kt, filt + 4 * ROWP4 + fslot);
                }
                // growth constants and row pointers of the epilogue
                if (!SINGLE) {
                    const float sg = __ldg(a.sigma + pidx), wt = __ldg(a.weights + pidx);
                    g_mu = __ldg(a.mu + pidx);
                    g_nc2 = -1.4426950408889634f / (2.f * sg * sg);
                    g_w2 = 2.f * wt;
                    g_wn = -wt;
                }
                const size_t dplane = (size_t)(b * a.C + j) * plane_elems;
                const int pitchB = a.pitch * 4;
                const long long o0B = (long long)(y0 + FT * it + F) * pitchB + (long long)((F + xw) * 4);
                EpiArgs e;
                e.u = reinterpret_cast<const float2 *>(filt) + tid;
                e.sa = reinterpret_cast<const char *>(a.state_in + dplane) + o0B;
                e.sb = e.sa + (long long)Lp * pitchB;
                e.oa = reinterpret_cast<char *>(a.out + dplane) + o0B;
                e.ob = e.oa + (long long)Lp * pitchB;
                e.pitchB = pitchB;
                e.pitch = a.pitch;
                e.H = a.H;
                e.W = a.W;
                e.ya0 = y0 + FT * it;
                e.Lp = Lp;
                e.xw = xw;
                e.xoff = xw < F ? a.W : (xw >= a.W - F ? -a.W : 0);
                e.row_wrap = a.row_wrap;
                e.mu = g_mu;
                e.nc2 = g_nc2;
                e.w2 = g_w2;
                e.wn = g_wn;
                e.dt = a.dt;
                e.peer_up = a.final_src && a.peer_up_out ? a.peer_up_out + dplane : nullptr;
                e.peer_dn = a.final_src && a.peer_dn_out ? a.peer_dn_out + dplane : nullptr;
                const bool euler = a.final_src && lenia;
                float2 st[FT];
#if MARCH_PREFETCH_EPI
                if (!pro && col_ok && euler) {
                    if (edge == 2) epilogue_prefetch<2>(e, st);
                    else epilogue_prefetch<0>(e, st);
                }
#endif
                if (!pro) __syncthreads();
                // ---- inverse transforms (groups 0-7) | forward transforms of the block needed next (groups 8-15, last destination);
                //      prologue rounds: forward transforms on all 16 groups
                if (!pro && grp < FT) {
                    inverse_row(filt + grp * ROWP4, tw, lane16);
                } else if (do_fwd) {
#if !MARCH_PREFETCH
                    forward_fetch(zf, vf);
#endif
                    forward_compute(vf, ring + (zf % RING) * ROWP4, tw, lane16);
                }
                if (pro) continue;
                __syncthreads();
#if !MARCH_PREFETCH_EPI
                if (col_ok && euler) {
                    if (edge == 2) epilogue_prefetch<2>(e, st);
                    else epilogue_prefetch<0>(e, st);
                }
#endif
                // ---- epilogue: growth + weighted source sum (+ Euler / clamp), sample n = tid of the 8 row pairs
                float2 uu[FT];
                if (EARLY_U) {
                    if (col_ok) {
#pragma unroll
                        for (int r = 0; r < FT; ++r) uu[r] = e.u[r * 2 * ROWP4];
                    }
                    __syncthreads();   // `filt` is free for the next FIR
                }
                if (col_ok) {
#define MARCH_EPI(ACC, EUL)                                   \
    do {                                                      \
        if (edge == 2) epilogue<2, ACC, EUL, EARLY_U>(e, st, uu);      \
        else if (edge == 1) epilogue<1, ACC, EUL, EARLY_U>(e, st, uu); \
        else epilogue<0, ACC, EUL, EARLY_U>(e, st, uu);                \
    } while (0)
                    if (SINGLE || !a.acc_in) {
                        if (euler) MARCH_EPI(false, true);
                        else MARCH_EPI(false, false);
                    } else {
                        if (euler) MARCH_EPI(true, true);
                        else MARCH_EPI(true, false);
                    }
#undef MARCH_EPI
                }
            }
        }
        // row slab: this band's share of the halo is in the neighbours' memory -> release it and bump their arrival counters
        if (a.final_src && (a.sig_up || a.sig_dn) && (top || bot)) {
            __threadfence_system();
            __syncthreads();
            if (tid == 0) {
                if (top && a.sig_up) asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(a.sig_up) : "memory");
                if (bot && a.sig_dn) asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(a.sig_dn) : "memory");
            }
        }
    }
#ifdef B200CA_DEV
    if (a.trace) {
        __syncthreads();
        if (tid == 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            a.trace[3 * blockIdx.x + 1] = t1;
        }
    }
#endif
}

// Tap table: kf[pair][dy][f] = (Khat[dy][k0(f)], Khat[dy][k1(f)]),  Khat[dy][k] = (1/N) sum_dx K[R+dy][R+dx] cos(2 pi k dx / N)
// with the FIR thread's two frequencies  ka = f >> 3, j = f & 7, k0 = ka + 32 j, k1 = k0 + 16.
__global__ void march_prepare_kernel(const float *__restrict__ K, float2 *__restrict__ kf, int nker, int ks, int *__restrict__ asym_flag) {
    const int ker = blockIdx.y;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= ROW4 || ker >= nker) return;
    const int ka = f >> 3, j = f & 7;
    const int k0 = ka + 32 * j, k1 = k0 + 16;
    const int off = (2 * R + 1 - ks) / 2;
    const float *kk = K + (size_t)ker * ks * ks;
    for (int dy = 0; dy < KROWS; ++dy) {
        double a0 = 0.0, a1 = 0.0;
        const int rp = R + dy - off, rm = R - dy - off;
        if (rp >= 0 && rp < ks && rm >= 0) {
            for (int c = 0; c < ks; ++c) {
                const int dx = c + off - R;
                const double up = kk[rp * ks + c], um = kk[rm * ks + c], mir = kk[rp * ks + (ks - 1 - c)];
                if (fabs(up - um) > 1e-6 * fmax(fabs(up), fabs(um)) + 1e-12 || fabs(up - mir) > 1e-6 * fmax(fabs(up), fabs(mir)) + 1e-12)
                    atomicExch(asym_flag, 1);
                const double s = 0.5 * (up + um);
                a0 += s * cospi(2.0 * (double)(((long long)k0 * dx % N + N) % N) / (double)N);
                a1 += s * cospi(2.0 * (double)(((long long)k1 * dx % N + N) % N) / (double)N);
            }
        }
        kf[((size_t)ker * KROWS + dy) * ROW4 + f] = make_float2((float)(a0 / N), (float)(a1 / N));
    }
}

// Launch plan of a balanced launch (host side; the kernel derives every CTA's bands from these numbers): grid size, and either
// split_k = 0 -- equal-cost runs of the (column, unit) space, see the kernel -- or every column cut into split_k equal bands,
// whichever leaves the busiest CTA with fewer rounds (a run pays PRO_UNITS rounds for every column it enters).
static void plan_balanced(int H, int W, int B, int waves, int *grid, int *split_k, int *units, int *strips, int *skew) {
    *units = ceil_div(H, 2 * FT);
    *strips = ceil_div(W, V);
    const long long ncols = (long long)*strips * B, total = *units * ncols;
    const long long G = std::min<long long>(total, 2LL * num_sms() * waves);
    *grid = (int)G;
    *split_k = 0;
    if (ncols <= G) {
        const long long k = std::min<long long>(G / ncols, *units);
        const long long rounds_split = ceil_div64(*units, k), rounds_runs = ceil_div64((*units + PRO_UNITS) * ncols, G);
        if (rounds_split <= rounds_runs) {
            *split_k = (int)k;
            *grid = (int)(k * ncols);
        }
    }
    // a full wave of two CTAs per SM, long enough for the skew to matter and (split launches) an even number of bands per column
    int sk = MARCH_SKEW;
#ifdef B200CA_DEV
    if (const char *e = getenv("B200CA_MARCH_SKEW")) sk = atoi(e);
#endif
    *skew = (waves == 1 && *grid == 2 * num_sms() && total >= 8LL * *grid && (*split_k == 0 || *split_k % 2 == 0)) ? sk : 0;
}

// The bands of CTA `cta` of a balanced launch, exactly as the kernel derives them (host copy: tests/test_host_logic.py checks that
// the bands of all CTAs tile every column once): writes up to `cap` (column, first unit, units) triples, returns their number.
static int bands_of_cta(int units, int ncols_strips_times_B, int grid, int split_k, int skew, int cta, int *out, int cap) {
    int n = 0, done = 0, col, u0, nu;
    while (next_band(units, ncols_strips_times_B, grid, split_k, skew, cta, done, col, u0, nu))
        if (n < cap) out[3 * n] = col, out[3 * n + 1] = u0, out[3 * n + 2] = nu, ++n;
    return n;
}

// row pairs per band: tall bands amortise the 32 extra forward transforms; small worlds trade that for enough CTAs
static int choose_lp(int H, int W, int B, int C) {
    const long long strips = ceil_div(W, V);
    const long long slots = 2LL * num_sms();
    if (MARCH_FORCELP) return MARCH_FORCELP;
    int best = FT;
    double best_cost = 1e30;
    for (int lp = MARCH_MAXLP; lp >= FT; lp >>= 1) {
        const long long ctas = strips * ceil_div(H, 2 * lp) * B;
        // Fitted to B200 timings of 1024^2 ... 16384^2 worlds with band heights 8 ... 256 (profiles/r02_conv_paths.md): a
        // launch behaves like (waves + 1.5) rounds of CTAs -- co-resident CTAs of a short launch run their phases in lockstep and
        // overlap less than the staggered CTAs of a long one -- and a CTA costs lp + 32 forward transforms plus lp x C x
        // (FIR + inverse + epilogue ~ 2.5 forward transforms each).
        const double waves = (double)ceil_div64(ctas, slots) + 1.5;
        const double cost = waves * ((lp + 32) + 2.5 * C * lp);
        if (cost < best_cost * 0.98) {   // near ties go to the taller band
            best_cost = cost;
            best = lp;
        }
    }
    return best;
}

}  // namespace march
}  // namespace b200ca

using namespace b200ca;

extern "C" int b200ca_lenia_march_supported(int H, int W) { return (H >= 2 * march::F && W >= 2 * march::F) ? 1 : 0; }

// conv_path = 'auto' (profiles/r02_conv_paths.md has the measurements behind the rule), by kernel size, sources and world size:
//   one source kernel per destination (C = 1): big worlds (more than 1536^2 cells in the launch) take the marching kernel
//     whatever the kernel size (4096^2: 95 us against 104 us for the direct kernel at k_size 5); smaller ones the direct
//     kernel up to k_size 11 (it evaluates the (k/2+1) x k folded footprint only) and the two-kernel hybrid above (a single
//     wave of CTAs: the shorter serial chain);
//   several sources: k_size <= 7 direct; k_size <= 11 direct for big worlds, hybrid for small ones; larger kernels (all of
//     demo_data is k_size 31): the marching kernel from 2560^2 cells up (C = 3, 4096^2: 784 us against 848 us), the hybrid below
//     (its fused kernel sums the sources in registers; the marching kernel accumulates through the output planes, one launch
//     per source: 140 us against 61 us at 1024^2);
//   whatever the preferred path cannot do falls to the next one, the direct convolution last.
extern "C" int b200ca_lenia_pick_algo_ks(int B, int NS, int H, int W, int ks) {
    const bool march_ok = b200ca_lenia_march_supported(H, W) != 0, rows_ok = b200ca_lenia_fft_length(H, W) > 0;
    const long long cells = (long long)B * H * W;
    const bool small = cells <= 1536LL * 1536LL;
    if (NS == 1) {
        if (!small && march_ok) return B200CA_ALGO_MARCH;
        if (ks <= 11) return B200CA_ALGO_DIRECT;
        return rows_ok ? B200CA_ALGO_FFT_ROWS : (march_ok ? B200CA_ALGO_MARCH : B200CA_ALGO_DIRECT);
    }
    if (ks <= 7) return B200CA_ALGO_DIRECT;
    if (ks <= 11 && !(small && rows_ok)) return B200CA_ALGO_DIRECT;
    if (cells >= 2560LL * 2560LL && march_ok && ks > 11) return B200CA_ALGO_MARCH;
    if (rows_ok) return B200CA_ALGO_FFT_ROWS;
    return march_ok ? B200CA_ALGO_MARCH : B200CA_ALGO_DIRECT;
}
extern "C" int b200ca_lenia_pick_algo(int B, int NS, int H, int W) { return b200ca_lenia_pick_algo_ks(B, NS, H, W, 31); }

extern "C" int64_t b200ca_lenia_march_kernel_elems(int B, int C, int k_mult) {
    return (int64_t)B * C * k_mult * C * march::KROWS * march::ROW4 * 2;
}

extern "C" int b200ca_lenia_march_band_rows(int H, int W, int B, int C) { return 2 * march::choose_lp(H, W, B, C); }

extern "C" int b200ca_lenia_march_plan(int H, int W, int B, int *grid, int *split_k, int *units, int *columns) {
    B200CA_REQUIRE(grid && split_k && units && columns && H >= 1 && W >= 1 && B >= 1, "lenia_march_plan: bad arguments");
    int strips = 0, skew = 0;
    march::plan_balanced(H, W, B, MARCH_JOB_WAVES > 0 ? MARCH_JOB_WAVES : 1, grid, split_k, units, &strips, &skew);
    *columns = strips * B;
    return 0;
}

extern "C" int b200ca_lenia_march_plan_bands(int H, int W, int B, int cta, int *bands, int cap) {
    int grid = 0, split_k = 0, units = 0, strips = 0, skew = 0;
    march::plan_balanced(H, W, B, MARCH_JOB_WAVES > 0 ? MARCH_JOB_WAVES : 1, &grid, &split_k, &units, &strips, &skew);
    if (!bands || cap < 1 || cta < 0 || cta >= grid) return -1;
    return march::bands_of_cta(units, strips * B, grid, split_k, skew, cta, bands, cap);
}

extern "C" int b200ca_lenia_march_prepare_kernel(const float *K, float *Kf, int B, int C, int k_mult, int ks, void *stream) {
    B200CA_REQUIRE(K && Kf && B >= 1 && C >= 1 && k_mult >= 1, "lenia_march_prepare_kernel: bad arguments");
    if (ks < 1 || ks > 2 * march::R + 1 || (ks % 2) == 0) {
        set_error("lenia_march_prepare_kernel: k_size %d not supported (odd, <= 31)", ks);
        return B200CA_E_UNSUPPORTED;
    }
    cudaStream_t st = as_stream(stream);
    int *flag = nullptr;
    B200CA_CUDA(cudaMalloc(&flag, sizeof(int)));
    int h_flag = 0;
    int rc = check_cuda(cudaMemsetAsync(flag, 0, sizeof(int), st), "memset");
    if (!rc) {
        const int nker = B * C * k_mult * C;
        dim3 grid(1, nker);
        march::march_prepare_kernel<<<grid, march::ROW4, 0, st>>>(K, reinterpret_cast<float2 *>(Kf), nker, ks, flag);
        rc = check_cuda(cudaPeekAtLastError(), "march_prepare_kernel");
    }
    if (!rc) rc = check_cuda(cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st), "D2H flag");
    if (!rc) rc = check_cuda(cudaStreamSynchronize(st), "sync");
    cudaFree(flag);
    if (rc) return rc;
    if (h_flag) {
        set_error("lenia_march_prepare_kernel: kernel rows are not symmetric (the FFT path needs K[dy][dx]==K[-dy][dx]==K[dy][-dx])");
        return B200CA_E_UNSUPPORTED;
    }
    return 0;
}

static int march_launch(march::Args a, int B, int C, int k_mult, int H, int W, int band_begin, int band_end, cudaStream_t st) {
    B200CA_REQUIRE(a.state_in && a.out && a.kf && a.mu && a.sigma && a.weights, "lenia_step_march: null pointer");
    B200CA_REQUIRE(a.state_in != a.out, "lenia_step_march: in-place update is not supported");
    B200CA_REQUIRE(a.mode == B200CA_MODE_LENIA || a.mode == B200CA_MODE_AFF, "lenia_step_march: bad mode %d", a.mode);
    B200CA_REQUIRE(B >= 1 && B <= 65535 && C >= 1 && k_mult >= 1, "lenia_step_march: bad B/C/k_mult");
    if (!b200ca_lenia_march_supported(H, W)) {
        set_error("lenia_step_march: world %dx%d not supported (H, W >= 32)", H, W);
        return B200CA_E_UNSUPPORTED;
    }
    static bool attr_done[64] = {false};
    int dev = 0;
    B200CA_CUDA(cudaGetDevice(&dev));
    if (!attr_done[dev & 63]) {
        B200CA_CUDA(cudaFuncSetAttribute(march::lenia_march_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)march::SMEM));
        B200CA_CUDA(cudaFuncSetAttribute(march::lenia_march_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)march::SMEM));
        attr_done[dev & 63] = true;
    }
    a.B = B;
    a.C = C;
    a.NS = C * k_mult;
    a.k_mult = k_mult;
    a.H = H;
    a.W = W;
    a.pitch = b200ca_frame_pitch(W);
    a.Lp = march::choose_lp(H, W, B, C);
    a.niter = a.Lp / march::FT;
    a.nbands = ceil_div(H, 2 * a.Lp);
    int job_waves = MARCH_JOB_WAVES;
#ifdef B200CA_DEV
    if (const char *e = getenv("B200CA_MARCH_WAVES")) job_waves = atoi(e);
    if (const char *e = getenv("B200CA_MARCH_TRACE")) a.trace = reinterpret_cast<unsigned long long *>(strtoull(e, nullptr, 0));   // device pointer, >= 3 x grid u64
#endif
    dim3 grid;
    if (band_begin == 0 && band_end == 0 && job_waves > 0) {
        // balanced launch: the whole world / slab, cut into equal runs of 16-row units for one wave of CTAs (see the kernel)
        a.jobs = 1;
        int g = 0;
        march::plan_balanced(H, W, B, job_waves, &g, &a.split_k, &a.units, &a.strips, &a.skew);
        grid = dim3((unsigned)g, 1, 1);
    } else {
        if (band_begin == 0 && band_end == 0) band_end = a.nbands;
        B200CA_REQUIRE(band_begin >= 0 && band_end <= a.nbands && band_begin <= band_end, "lenia_step_march: band range [%d,%d) outside [0,%d)",
                       band_begin, band_end, a.nbands);
        if (band_begin == band_end) return 0;
        a.band0 = band_begin;
        grid = dim3(ceil_div(W, march::V), band_end - band_begin, B);
    }
    for (int s = 0; s < a.NS; ++s) {
        a.src = s;
        a.acc_in = s > 0;
        a.final_src = s == a.NS - 1;
        if (C == 1) B200CA_CUDA(launch_pdl(march::lenia_march_kernel<true>, grid, dim3(march::THREADS), march::SMEM, st, a));
        else B200CA_CUDA(launch_pdl(march::lenia_march_kernel<false>, grid, dim3(march::THREADS), march::SMEM, st, a));
    }
    B200CA_LAUNCH_CHECK("lenia_march_kernel");
    if (a.jobs) note_variant("lenia_march<%s,balanced:%u%s%s>x%d", C == 1 ? "single" : "multi", grid.x, a.split_k ? ",split" : "", (a.sig_up || a.sig_dn) ? ",p2p-halo" : "", a.NS);
    else note_variant("lenia_march<%s,Lp=%d%s>x%d", C == 1 ? "single" : "multi", a.Lp, (a.sig_up || a.sig_dn) ? ",p2p-halo" : "", a.NS);
    return 0;
}

extern "C" int b200ca_lenia_step_march(const float *state_in, float *state_out, const float *Kf, const float *mu, const float *sigma,
                                       const float *weights, int B, int C, int k_mult, int H, int W, float dt, int mode, int row_wrap,
                                       int band_begin, int band_end, void *stream) {
    march::Args a = {};
    a.state_in = state_in;
    a.out = state_out;
    a.kf = reinterpret_cast<const float2 *>(Kf);
    a.mu = mu;
    a.sigma = sigma;
    a.weights = weights;
    a.dt = dt;
    a.mode = mode;
    a.row_wrap = row_wrap;
    return march_launch(a, B, C, k_mult, H, W, band_begin, band_end, as_stream(stream));
}

extern "C" int b200ca_lenia_march_slab_ctas(int H, int W, int B, int C) {
    (void)H;
    return ceil_div(W, march::V) * B;
}

extern "C" int b200ca_lenia_step_march_slab(const float *state_in, float *state_out, const float *Kf, const float *mu, const float *sigma,
                                            const float *weights, int B, int C, int k_mult, int H, int W, float dt, int mode,
                                            float *peer_up_out, float *peer_dn_out, uint32_t *sig_up, uint32_t *sig_dn,
                                            const uint32_t *wait_top, const uint32_t *wait_bot, uint32_t wait_target, uint32_t *err_flag,
                                            void *stream) {
    B200CA_REQUIRE(peer_up_out && peer_dn_out && sig_up && sig_dn && wait_top && wait_bot, "lenia_step_march_slab: null peer pointer");
    march::Args a = {};
    a.state_in = state_in;
    a.out = state_out;
    a.kf = reinterpret_cast<const float2 *>(Kf);
    a.mu = mu;
    a.sigma = sigma;
    a.weights = weights;
    a.dt = dt;
    a.mode = mode;
    a.row_wrap = 0;
    a.peer_up_out = peer_up_out;
    a.peer_dn_out = peer_dn_out;
    a.sig_up = sig_up;
    a.sig_dn = sig_dn;
    a.wait_top = wait_top;
    a.wait_bot = wait_bot;
    a.wait_target = wait_target;
    a.err_flag = err_flag;
    return march_launch(a, B, C, k_mult, H, W, 0, 0, as_stream(stream));
}
