// Shared definitions of the NCA kernels (nca.cu: SIMT path + life-mask pass; nca_tc.cu: tcgen05 path).
#pragma once
#include "common.cuh"

namespace b200ca {

constexpr int NCA_N = 16;     // state channels (all demo_data/NCA_models checkpoints)
constexpr int NCA_HID = 128;  // hidden units
constexpr int NCA_K1 = 32;    // folded perception features: 8 sobel-x + 8 sobel-y + 16 identity

// Wprep layout.  SIMT section (floats, offset 0):
//   W1f[128][32]  folded first layer: col f<8: W1[:,2f]+W1[:,2f+1] (sobel-x of channel f, nca.py:341 quirk:
//                 conv output o reads channel o//2); col 8+g: W1[:,16+2g]+W1[:,17+2g] (sobel-y of channel 8+g);
//                 col 16+c: W1[:,32+c] (identity)
//   b1[128], W2t[128][16] (W2 transposed), b2[16]
constexpr int WP_W1F = 0;
constexpr int WP_B1 = WP_W1F + NCA_HID * NCA_K1;
constexpr int WP_W2T = WP_B1 + NCA_HID;
constexpr int WP_B2 = WP_W2T + NCA_HID * NCA_N;
constexpr int WP_SIMT_FLOATS = WP_B2 + NCA_N;  // 6288
// Tensor-core section (byte offset WP_TC_OFF): UMMA-ready operand images, see nca_tc.cu
constexpr int WP_TC_OFF = 32 * 1024;
constexpr int WP_TC_BYTES = 96 * 1024;
constexpr int WP_TOTAL_BYTES = WP_TC_OFF + WP_TC_BYTES;

struct NcaArgs {
    const float *in;
    float *out;
    const float *wp;       // Wprep
    const uint8_t *mask;   // (B,H,W) or null
    uint32_t *alive_bits;  // [B*H*wpr] post-update alpha > 0.1
    uint32_t *nz_bits;     // [B*H*wpr] cell written non-zero by the update pass
    uint64_t seed, step;
    uint64_t cell0;        // global index of this launch's first cell (batch shards draw their slice of ONE global mask)
    int B, H, W, wpr;
};

int nca_step_tc(const NcaArgs &a, cudaStream_t st);
int nca_prepare_tc(const float *W1, const float *b1, const float *W2, const float *b2, void *Wprep, cudaStream_t st);

// update decision of one cell: explicit mask byte, or Philox4x32-10(seed; cell, step) -> u <= 0.5
__device__ __forceinline__ bool nca_update_bit(const NcaArgs &a, size_t cell) {
    if (a.mask) return a.mask[cell] != 0;
    uint32_t r[4];
    Philox::gen(a.seed, (uint64_t)cell + a.cell0, a.step, r);
    return (r[0] >> 8) <= (1u << 23);  // u = (r>>8) * 2^-24 <= 0.5  (nca.py:256 `rand <= 0.5`)
}

}  // namespace b200ca
