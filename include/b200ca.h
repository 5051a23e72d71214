/*
 * b200ca.h -- C ABI of the B200-native (sm_100a) cellular-automaton step kernels.
 *
 * The reference (frotaur/PyCA) is pure Python + PyTorch and has NO foreign-function
 * interface: its hot path is `Automaton.step()` of the model classes in
 * `pyca/automata/models/`.  The entry points below are therefore what a binding
 * for that path replaces, one per reference method (file:line relative to the PyCA
 * tree).  The reference-side binding (a ctypes stub inside each `step()`) is shown
 * in INTEGRATION.md; `pyca_b200/_lib.py` is that stub for this repository.
 *
 * Conventions
 *   - `extern "C"`, plain pointers and sizes; no torch types.
 *   - Pointers without the `h_` prefix are DEVICE pointers owned by the caller, valid on
 *     the current CUDA device; `h_` pointers are HOST buffers (the `*_host` entry points
 *     do their own H2D/D2H copies and are what `e2e` in bench.py times).
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream).
 *     Every call only enqueues work on that stream unless documented otherwise.
 *   - Return value: 0 on success; >0 a `cudaError_t`; <0 a B200CA_E_* code.  The text
 *     of the last error on the calling thread is `b200ca_last_error()`.
 *   - No allocation per call: the library keeps small per-device caches (TMA tensor maps,
 *     scratch) released by `b200ca_shutdown()`.
 *   - There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef B200CA_H
#define B200CA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200CA_VERSION 100

#define B200CA_E_BADARG (-1)   /* inconsistent sizes / null pointers / unsupported shape */
#define B200CA_E_NODEVICE (-2) /* no sm_100 CUDA device */
#define B200CA_E_DRIVER (-3)   /* driver entry point (TMA descriptor encode) unavailable/failed */
#define B200CA_E_UNSUPPORTED (-4)

int b200ca_version(void);
const char *b200ca_last_error(void);
/* Test hook: ';'-separated names of the kernel instantiations the compute calls of this thread launched since the
 * last reset (e.g. "lenia_fft_rows<1,1024,split>;lenia_fft_firinv<8,single,1,512>"); reset != 0 clears the log after
 * reading it.  The parity tests use it to assert which template instantiation a shape dispatched to. */
const char *b200ca_debug_variants(int reset);
/* sm count / compute capability of the current device (error if none). */
int b200ca_device_info(int *sm_count, int *cc_major, int *cc_minor);
/* frees cached descriptors/scratch of all devices. */
int b200ca_shutdown(void);

/* ------------------------------------------------------------------------------------
 * Game of Life / outer-totalistic CA2D.   Replaces CA2D.step  (models/ca2d.py:195-209,
 * get_nth_bit :221-225).  The world lives bit-packed: 32 cells per uint32, cell x of row y
 * is bit (x & 31) of word  y*stride_words + (x >> 5);  bits at x >= W are kept zero.
 * ---------------------------------------------------------------------------------- */
int b200ca_gol_words_per_row(int W); /* ceil(W/32) */

/* elem_bytes in {1,2,4,8}: uint8/bool, int16 (world after reset(), ca2d.py:227-246), int32 (world
 * after any step, ca2d.py:209), int64.  A cell is alive iff its value == 1 (ca2d.py:208). */
int b200ca_gol_pack(const void *world, int elem_bytes, uint32_t *bits, int H, int W, int stride_words, void *stream);
int b200ca_gol_unpack(const uint32_t *bits, void *world, int elem_bytes, int H, int W, int stride_words, void *stream);

/* One step.  s_mask/b_mask: 9-bit survive/birth masks, bit n set = rule holds for n live
 * neighbours (ca2d.py:80-88; Life = S 0b1100, B 0b1000).  vert_open = 0: torus (reference
 * semantics);  1: rows outside [0,H) are dead (used on a row slab that carries ghost rows;
 * columns always wrap).  bits_in != bits_out. */
int b200ca_gol_step(const uint32_t *bits_in, uint32_t *bits_out, int H, int W, int stride_words, uint32_t s_mask,
                    uint32_t b_mask, int vert_open, void *stream);
/* nsteps steps ping-ponging a<->b; the result is in `a` if nsteps is even, else in `b`. */
int b200ca_gol_run(uint32_t *a, uint32_t *b, int H, int W, int stride_words, uint32_t s_mask, uint32_t b_mask,
                   int vert_open, int nsteps, void *stream);
/* population count of the packed world into *count (device, uint64). */
int b200ca_gol_population(const uint32_t *bits, int H, int W, int stride_words, unsigned long long *count, void *stream);
/* fills the packed world with Bernoulli(1/2) bits from a counter-based generator (synthetic
 * worlds too large to build as int32).  row0 offsets the row counter so
 * slabs of one world can be generated independently. */
int b200ca_gol_random_fill(uint32_t *bits, int H, int W, int stride_words, uint64_t seed, int64_t row0, void *stream);
/* HOST world in, HOST world out (int32 0/1, H*W each), nsteps steps; copies inside. */
int b200ca_gol_run_host(const int32_t *h_world_in, int32_t *h_world_out, int H, int W, uint32_t s_mask, uint32_t b_mask,
                        int nsteps);

/* ------------------------------------------------------------------------------------
 * Lenia family.  State planes live in a FRAMED layout so the periodic wrap (and a row-slab
 * neighbour's rows) are plain reads:  plane (b,c) has (H + 2*B200CA_FRAME) rows of `pitch`
 * floats; cell (y,x) is at  ((b*C + c)*(H+2F) + y + F)*pitch + x + F.  The frame holds the
 * wrapped copies of the opposite edges; the step kernels keep it current for their output.
 * ---------------------------------------------------------------------------------- */
#define B200CA_FRAME 16
int b200ca_frame_pitch(int W);                 /* floats per framed row: roundup(W + 2F, 4) */
int64_t b200ca_frame_elems(int B, int C, int H, int W); /* floats in a framed (B,C,H,W) tensor */
/* dense (B,C,H,W) -> framed (interior + wrapped frame) and back.  `wrap_rows` = 0 leaves the top/bottom
 * frame rows untouched (row slab: they are ghost rows owned by the neighbours). */
int b200ca_frame_load(const float *dense, float *framed, int B, int C, int H, int W, int wrap_rows, void *stream);
int b200ca_frame_store(const float *framed, float *dense, int B, int C, int H, int W, void *stream);
/* re-derives the frame of `framed` from its interior (after in-place edits of the interior). */
int b200ca_frame_refresh(float *framed, int B, int C, int H, int W, int wrap_rows, void *stream);

/* Kernel preparation (cold; called from update_params, lenia.py:143-205).  K is the reference's
 * `self.k` (B, C*k_mult, C, ks, ks) fp32 (lenia.py:202, :314-362), ks odd <= 31.  Kprep receives
 * b200ca_lenia_kprep_elems() floats: the D4-symmetric kernel centred in 31x31 and folded to the
 * rows dy >= 0.  Returns B200CA_E_UNSUPPORTED if K is not symmetric under row reversal. */
int64_t b200ca_lenia_kprep_elems(int B, int C, int k_mult);
int b200ca_lenia_prepare_kernel(const float *K, float *Kprep, int B, int C, int k_mult, int ks, void *stream);

#define B200CA_MODE_LENIA 0 /* out = clamp(state + dt * dx, 0, 1)            lenia.py:430-451        */
#define B200CA_MODE_AFF 1   /* out = dx (affinity, framed with pad>=2)        macelenia.py:122-159    */

/* Direct (spatial) convolution + growth + weighted source sum (+ Euler/clamp).
 *   U[b,i*km+m,j] = K[b,i*km+m,j] (*) state[b,i]   (periodic; lenia.py:453-469 computes it by FFT)
 *   G = 2*exp(-((U-mu)^2/sigma^2)/2) - 1           (lenia.py:423-426)
 *   dx[b,j] = sum_{i,m} weights[b,i*km+m,j] * G    (lenia.py:444-446)
 * state_in/state_out: framed (B,C,H,W); mu/sigma/weights: (B, C*k_mult, C) fp32 device.
 * row_wrap = 1: the kernel also maintains the top/bottom frame of `out` (torus);
 *            0: top/bottom frame rows of `out` are left to the caller (row slab / halo exchange).
 * tile_row_begin/end: restricts the launch to output tile-rows [begin,end) in units of
 *            b200ca_lenia_tile_rows() rows (0,0 = all) so slab boundaries can be issued first. */
int b200ca_lenia_tile_rows(int H, int W, int B, int C);
int b200ca_lenia_step(const float *state_in, float *state_out, const float *Kprep, const float *mu, const float *sigma,
                      const float *weights, int B, int C, int k_mult, int H, int W, float dt, int mode, int row_wrap,
                      int tile_row_begin, int tile_row_end, void *stream);

/* The same step with `g_arbi` growth functions (lenia.py:390-418 -> ArbitraryFunction.forward, funcgen.py:79-129): the growth of
 * kernel pair p is the Fourier series  sum_h g_cos[p,h] cos(g_freq[p,h] u) + g_sin[p,h] sin(g_freq[p,h] u)  with
 * g_freq = fp32(2 pi / period) * harmonic (period 2 for growth functions); rescale_on: followed by the reference's global
 * rescale (v - min) / (max - min + 1e-8) * (hi - lo) + lo with min/max over the whole (H,W) plane of each pair (costs a
 * second pass of the convolution; g_stats = 2 floats per pair of device scratch); clip_on: then max(v, clip_min).
 * g_cos/g_sin/g_freq: (B, C*k_mult, C, n_harmonics) fp32 device.  Direct path only. */
int b200ca_lenia_step_garbi(const float *state_in, float *state_out, const float *Kprep, const float *weights, const float *g_cos,
                            const float *g_sin, const float *g_freq, int n_harmonics, int rescale_on, float rescale_lo,
                            float rescale_hi, int clip_on, float clip_min, float *g_stats, int B, int C, int k_mult, int H, int W,
                            float dt, int mode, int row_wrap, void *stream);

/* FFT path of the same operator (row FFT x column-direct hybrid, see pyca_b200/csrc/lenia_fft.cu): identical
 * arguments and semantics to b200ca_lenia_step; supported when H is even (>= 32) and W >= 64.  The x direction is an
 * exact periodic transform when W is 256, 1024 or 4096, overlap-save segments otherwise.
 *   b200ca_lenia_fft_length        transform length N the path would use for an HxW world (0 = unsupported)
 *   b200ca_lenia_fft_kernel_elems  floats of the prepared kernel spectra  (B, C*k_mult, C, 16, N)
 *   b200ca_lenia_fft_workspace_elems floats of the row-spectra workspace the step needs
 *   b200ca_lenia_fft_prepare_kernel  K (B, C*k_mult, C, ks, ks) -> Kfft (cold path, update_params) */
int b200ca_lenia_fft_length(int H, int W);
int64_t b200ca_lenia_fft_kernel_elems(int B, int C, int k_mult, int H, int W);
int64_t b200ca_lenia_fft_workspace_elems(int B, int C, int k_mult, int H, int W);
int b200ca_lenia_fft_prepare_kernel(const float *K, float *Kfft, int B, int C, int k_mult, int ks, int H, int W, void *stream);
int b200ca_lenia_step_fft(const float *state_in, float *state_out, const float *Kfft, const float *mu, const float *sigma,
                          const float *weights, float *workspace, int B, int C, int k_mult, int H, int W, float dt, int mode,
                          int row_wrap, void *stream);
/* The same step in two parts, for a row slab that overlaps its halo exchange with compute: phase 1 = row FFTs of the
 * spectra rows built from owned rows only (needs no ghost rows), phase 2 = the remaining 30 row FFTs + FIR + inverse
 * (after the ghost rows arrived); phase 0 = everything (what b200ca_lenia_step_fft does). */
int b200ca_lenia_step_fft_phase(const float *state_in, float *state_out, const float *Kfft, const float *mu, const float *sigma,
                                const float *weights, float *workspace, int B, int C, int k_mult, int H, int W, float dt,
                                int mode, int row_wrap, int phase, void *stream);

/* "Marching" FFT path (pyca_b200/csrc/lenia_march.cu), the default for every world with H, W >= 32: the same operator and
 * arguments as b200ca_lenia_step, computed by ONE kernel per source channel in which a CTA marches down a band of rows of
 * one 224-column overlap-save strip -- forward row transforms (N = 256), the real 31-tap column FIR on the spectra, inverse
 * transforms, growth, source sum and Euler step -- with the spectra kept in a shared-memory ring (no workspace, no HBM
 * round trip of the spectra).
 *   b200ca_lenia_march_supported      1 if the path handles an HxW world
 *   b200ca_lenia_march_kernel_elems   floats of the prepared tap table (B, C*k_mult, C, 16, 256)
 *   b200ca_lenia_march_prepare_kernel K (B, C*k_mult, C, ks, ks) -> tap table (cold path, update_params)
 *   b200ca_lenia_march_band_rows      image rows per band for this shape; band_begin/band_end of the step restrict a launch
 *                                     to bands [begin, end) (0,0 = all) so that a row slab can issue its boundary bands first */
#define B200CA_ALGO_DIRECT 0
#define B200CA_ALGO_FFT_ROWS 1
#define B200CA_ALGO_MARCH 2
/* which of the three convolution paths conv_path='auto' takes for B worlds of HxW with NS = C*k_mult source kernels per
 * destination plane (a rule fitted to B200 measurements; Gaussian growth -- Fourier growth functions are direct-only) */
int b200ca_lenia_pick_algo(int B, int NS, int H, int W);                 /* k_size 31 */
int b200ca_lenia_pick_algo_ks(int B, int NS, int H, int W, int k_size);  /* per kernel size (lenia.py:143-168 k_size_override) */
int b200ca_lenia_march_supported(int H, int W);
int64_t b200ca_lenia_march_kernel_elems(int B, int C, int k_mult);
int b200ca_lenia_march_prepare_kernel(const float *K, float *Kf, int B, int C, int k_mult, int ks, void *stream);
int b200ca_lenia_march_band_rows(int H, int W, int B, int C);
/* Host-side copy of the plan of a whole-world launch (band_begin = band_end = 0), which is BALANCED: one wave of `grid` CTAs shares
 * the `columns` = strips x B columns of bands, each `units` 16-row units tall, either in equal-cost runs (split_k = 0) or with
 * every column cut into split_k equal bands.  b200ca_lenia_march_plan_bands writes the (column, first unit, units) triples of
 * one CTA's bands exactly as the kernel derives them and returns their number (-1: bad arguments).  No device needed. */
int b200ca_lenia_march_plan(int H, int W, int B, int *grid, int *split_k, int *units, int *columns);
int b200ca_lenia_march_plan_bands(int H, int W, int B, int cta, int *bands, int cap);
int b200ca_lenia_step_march(const float *state_in, float *state_out, const float *Kf, const float *mu, const float *sigma,
                            const float *weights, int B, int C, int k_mult, int H, int W, float dt, int mode, int row_wrap,
                            int band_begin, int band_end, void *stream);

/* Row slab of a torus sharded over the GPUs of one box, halo exchange by P2P stores from inside the step kernel (one launch per
 * step and source channel, no host-side send/recv): the slab's first / last 16 rows are also stored into the ghost rows of the
 * neighbours' `state_out` buffers (`peer_up_out` / `peer_dn_out`: those buffers mapped into this process, e.g. through
 * torch.distributed._symmetric_memory or CUDA IPC), then every boundary CTA adds 1 to the neighbour's arrival counter
 * (`sig_up`: the up neighbour's bottom-ghost counter, `sig_dn`: the down neighbour's top-ghost counter).  The CTAs of the first
 * / last band wait until this rank's own counters `wait_top` / `wait_bot` reach `wait_target` = step_index *
 * b200ca_lenia_march_slab_ctas() before they read their ghost rows (step 0 reads ghost rows the caller exchanged itself).
 * A wait that lasts longer than ~4 s sets *err_flag and goes on (no hang).  Every rank must issue the same sequence of steps. */
int b200ca_lenia_march_slab_ctas(int H, int W, int B, int C);
int b200ca_lenia_step_march_slab(const float *state_in, float *state_out, const float *Kf, const float *mu, const float *sigma,
                                 const float *weights, int B, int C, int k_mult, int H, int W, float dt, int mode,
                                 float *peer_up_out, float *peer_dn_out, uint32_t *sig_up, uint32_t *sig_dn,
                                 const uint32_t *wait_top, const uint32_t *wait_bot, uint32_t wait_target, uint32_t *err_flag,
                                 void *stream);

/* One halo exchange of a row slab as one kernel per rank (csrc/halo.cu): copies `nbytes` (multiple of 16) from src_up / src_dn
 * (this rank's first / last owned rows) into dst_up / dst_dn (the up neighbour's bottom ghost rows / the down neighbour's top
 * ghost rows, peer-mapped) and synchronises with both neighbours through counters in peer-mapped memory: `up_ready` / `dn_ready`
 * and `up_arrived` / `dn_arrived` are slots in the NEIGHBOURS' counter blocks that this rank bumps, `my_ready[2]` / `my_arrived[2]`
 * this rank's own ([0] written by the up neighbour, [1] by the down neighbour); `epoch` = 1, 2, ... counts the exchanges, every
 * rank issues the same sequence.  When the kernel ends this rank's ghost rows are current.  `ticket`: 4 zeroed bytes of local
 * scratch; a wait longer than ~4 s sets *err_flag. */
int b200ca_halo_push(const void *src_up, void *dst_up, const void *src_dn, void *dst_dn, int64_t nbytes, uint32_t *up_ready,
                     uint32_t *dn_ready, const uint32_t *my_ready, uint32_t *up_arrived, uint32_t *dn_arrived,
                     const uint32_t *my_arrived, uint32_t epoch, uint32_t *ticket, uint32_t *err_flag, void *stream);
/* The same for `nplanes` row blocks per direction, `plane_stride_bytes` apart in every buffer (the (b, c) planes of a framed Lenia /
 * MaCE slab: the MaCE slabs exchange affinity rows and state rows this way, slab.py: MaceSlabP2P). */
int b200ca_halo_push_planes(const void *src_up, void *dst_up, const void *src_dn, void *dst_dn, int64_t nbytes, int nplanes,
                            int64_t plane_stride_bytes, uint32_t *up_ready, uint32_t *dn_ready, const uint32_t *my_ready,
                            uint32_t *up_arrived, uint32_t *dn_arrived, const uint32_t *my_arrived, uint32_t epoch, uint32_t *ticket,
                            uint32_t *err_flag, void *stream);

/* Mass-conserving redistribution (+ optional cross-channel mixing).  Replaces
 * MaCELenia._mace_step (macelenia.py:89-120) after the affinity, and
 * MaCELeniaXChan._cross_chan_step (mace_lenia_xchan.py:68-76) when alpha_on != 0:
 *   E = exp(beta*Aff);  Z = sum3x3(E);  A' = E * sum3x3(A / Z);
 *   P = softmax_c(beta*Aff);  A'' = A' - alpha * (A' - (sum_c A') * P).
 * aff: framed (B,C,H,W) written by b200ca_lenia_step(mode=AFF) (frame valid to 2 cells). */
int b200ca_mace_redistribute(const float *state_in, const float *aff, float *state_out, int B, int C, int H, int W,
                             float beta, int alpha_on, float alpha, int row_wrap, void *stream);

/* MaCE food subsystem (macelenia.py:122-226; off by default in the reference).
 *   b200ca_lenia_conv_sum    out[b,j] = sum_{i,m} K[b,i*k_mult+m,j] (*) planes_in[b,i]   (framed in/out, direct path): the
 *                            `kernel_fftconv(food_aff).sum(dim=1)` of food sensing (:152-154) with planes_in = the food mask
 *                            replicated over the C channels.
 *   b200ca_mace_food_sense   aff += (food > 0) + food_conv over the whole framed extent (:149-157); food dense (B,1,H,W).
 *   b200ca_mace_food_decay   state -= (state > threshold ? decay : 0) on the interior, in place; loss[b] += mass removed
 *                            (double, device) (:171-185).  The frame of `state` is stale afterwards.
 *   b200ca_mace_food_consume transfer = min(food, rate * [food > 0 and sum_c state >= min_density]); state_out = state_in +
 *                            transfer / 3 (framed, periodic frame written); food -= transfer (in place) (:194-226, death off);
 *                            alpha_on: followed by the cross-channel mixing with softmax_c(beta * aff) that
 *                            MaCELeniaXChan.step applies after the food step (mace_lenia_xchan.py:59-76). */
int b200ca_lenia_conv_sum(const float *planes_in, float *out, const float *Kprep, int B, int C, int k_mult, int H, int W,
                          void *stream);
int b200ca_mace_food_sense(float *aff, const float *food_conv, const float *food, int B, int C, int H, int W, void *stream);
int b200ca_mace_food_decay(float *state, double *loss, int B, int C, int H, int W, float threshold, float decay, void *stream);
int b200ca_mace_food_consume(const float *state_in, float *state_out, float *food, const float *aff, int B, int C, int H, int W,
                             float min_density, float transfer_rate, int alpha_on, float beta, float alpha, void *stream);

/* FlowLenia step tail (flow_lenia.py:207-253): flow field F = grad(Aff)(1 - alpha) - grad(sum_c state) alpha with zero-padded
 * Sobel stencils and alpha = clip((sum_c state / theta_x)^n, 0, 1), then reintegration tracking (ReintegrationTracker,
 * :97-153): every cell's mass, a square of half-width sigma_rt displaced by clip(dt F, +-(dd - sigma_rt)), is collected by
 * the cells it overlaps within +-dd.  aff: framed output of the convolution in mode AFF.  state_out gets its periodic frame.
 * Mass conserving away from the world edge; like the reference, nothing flows across the edge. */
int b200ca_flow_reintegrate(const float *state_in, const float *aff, float *state_out, int B, int C, int H, int W, float dt, int dd,
                            float sigma_rt, float theta_x, int n_pow, void *stream);

/* sum over H*W of every (b,c) plane of a framed tensor into mass[b*C+c] (double, device): total_mass()
 * (macelenia.py:421-428). */
int b200ca_frame_mass(const float *framed, double *mass, int B, int C, int H, int W, void *stream);

/* HOST dense state in/out, device parameters prepared by the caller; nsteps Lenia (mode 0) or
 * MaCE (mode 1) / MaCE-XChan (mode 2) steps with the copies inside the call. */
int b200ca_lenia_run_host(const float *h_state_in, float *h_state_out, const float *Kprep, const float *mu,
                          const float *sigma, const float *weights, int B, int C, int k_mult, int H, int W, float dt,
                          int family, float beta, float alpha, int nsteps);

/* Pipelined host entry point: a stream of HOST-resident worlds (pinned buffers) through the same step.  The reference's
 * headless loop on host data is `auto.state = t; auto.step(); t = auto.state` per world (lenia.py:79, :430-451), every
 * copy serialised with the step; independent worlds can overlap instead.  A pipe owns `depth` lanes (stream, staging,
 * framed buffers, FFT workspace); submit() queues H2D -> nsteps steps -> D2H on the next lane (round-robin) and returns
 * at once -- it only blocks while that lane's previous world is still in flight; drain() waits for everything queued.
 * K/mu/sigma/weights: the reference's `k`, `mu`, `sigma`, `weights` tensors (host or device pointers; copied at create).
 * family: 0 Lenia, 1 MaCE, 2 MaCE-XChan.  Uses the FFT path when the shape supports it, else the direct path. */
int b200ca_lenia_pipe_create(void **pipe, const float *K, const float *mu, const float *sigma, const float *weights, int B,
                             int C, int k_mult, int ks, int H, int W, float dt, int family, float beta, float alpha, int depth);
int b200ca_lenia_pipe_submit(void *pipe, const float *h_state_in, float *h_state_out, int nsteps);
int b200ca_lenia_pipe_drain(void *pipe);
int b200ca_lenia_pipe_destroy(void *pipe);

/* ------------------------------------------------------------------------------------
 * Neural CA inference.  Replaces NCAModule.forward for one step (models/nca.py:239-263):
 * SobelConv perception (:328-344), Linear-ReLU-Linear (:205-208), sigmoid * update mask,
 * life masks (:226-237), clamp.  state: dense (B,n,H,W) fp32 NCHW, n == 16, hidden == 128.
 * ---------------------------------------------------------------------------------- */
/* Weight preparation (cold; after load_model, nca.py:53-64).  W1 (hid, 3n), b1 (hid), W2 (n, hid), b2 (n)
 * as in the checkpoint's state_dict.  Wprep gets b200ca_nca_wprep_bytes() bytes (folded W1 columns,
 * split-precision copies for the tensor-core path). */
int64_t b200ca_nca_wprep_bytes(int n, int hid);
int b200ca_nca_prepare_weights(const float *W1, const float *b1, const float *W2, const float *b2, void *Wprep, int n,
                               int hid, void *stream);
/* scratch for the post-update alive bits: bytes needed for (B,H,W). */
int64_t b200ca_nca_scratch_bytes(int B, int H, int W);

#define B200CA_NCA_SIMT 0  /* fp32 FMA MLP (reference-grade arithmetic, slow)             */
#define B200CA_NCA_TC 1    /* tcgen05 / TMEM MLP, split-tf32 (fp32-equivalent accuracy)    */

/* update_mask: (B,H,W) uint8, non-zero = cell updates (the reference's `torch.rand(B,1,H,W) <= 0.5`,
 * nca.py:256);  NULL = draw it in-kernel from Philox4x32-10(seed, step, cell).  */
int b200ca_nca_step(const float *state_in, float *state_out, const void *Wprep, const uint8_t *update_mask,
                    uint64_t seed, uint64_t step_index, void *scratch, int B, int n, int hid, int H, int W, int impl,
                    void *stream);
/* The same step for a batch SHARD: worlds [first_world, first_world + B) of a larger batch.  The in-kernel Philox counter is
 * the GLOBAL cell index, so the shards of a batch split over GPUs draw exactly the mask the unsplit batch draws
 * ("give each shard its slice of the same global mask"). */
int b200ca_nca_step_shard(const float *state_in, float *state_out, const void *Wprep, const uint8_t *update_mask, uint64_t seed,
                          uint64_t step_index, uint64_t first_world, void *scratch, int B, int n, int hid, int H, int W, int impl,
                          void *stream);
int b200ca_nca_run_host(const float *h_state_in, float *h_state_out, const void *Wprep, const uint8_t *h_update_mask,
                        uint64_t seed, int B, int n, int hid, int H, int W, int impl, int nsteps);

/* ------------------------------------------------------------------------------------
 * Fused draw(): state -> RGB frame on the device (the step either side of the hot path).
 * Replaces the torch elementwise ops of draw() (ca2d.py:117-124, lenia.py:481-500, macelenia.py:365-385,
 * nca.py:74-87) and the float permute + D2H + host uint8 cast of Automaton.worldmap (automaton.py:209-216).
 *   worldmap: float (3,H,W) device tensor = the reference's `_worldmap` (NULL = not wanted, except CA2D where it is the
 *             echo state read and rewritten every frame);
 *   frame:    uint8 (W,H,3) device buffer = what `worldmap` returns after its host-side cast:
 *             frame[x][y][c] = (uint8)(255.f * worldmap[c][y][x]) (truncation); NULL = not wanted.
 * ---------------------------------------------------------------------------------- */
int b200ca_draw_gol(const uint32_t *bits, float *worldmap, uint8_t *frame, int H, int W, int stride_words, float decay_r,
                    float decay_g, float decay_b, void *stream);
/* state: framed (B,C,H,W); C==1 grey, C==2 (c0,c1,0), C>=3 first three channels; food: dense (H,W) plane added to all
 * three colours before the clamp (macelenia.py:379-380) or NULL. */
int b200ca_draw_lenia(const float *state, const float *food, float *worldmap, uint8_t *frame, int batch_index, int C, int H,
                      int W, void *stream);
/* state: dense (B,n,H,W), channels 0..3 = RGBA; rgb = clamp(rgb,0,1) * clamp(a,0,1) (brush overlay is GUI, not drawn). */
int b200ca_draw_nca(const float *state, float *worldmap, uint8_t *frame, int batch_index, int n, int H, int W, void *stream);
/* Automaton.worldmap alone (automaton.py:216) for a float (3,H,W) `_worldmap` produced by any draw(): -> (W,H,3) uint8. */
int b200ca_frame_pack_rgb(const float *worldmap, uint8_t *frame, int H, int W, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200CA_H */
