// tcgen05 / TMEM implementation of the NCA update pass (placeholder until the kernel lands).
#include "nca.cuh"

namespace b200ca {

int nca_prepare_tc(const float *, const float *, const float *, const float *, void *, cudaStream_t) { return 0; }

int nca_step_tc(const NcaArgs &, cudaStream_t) {
    set_error("nca_step: tensor-core implementation not built yet");
    return B200CA_E_UNSUPPORTED;
}

}  // namespace b200ca
