// Tuned sm_100a propagation kernel for the standard dilation set (placeholder: not selected yet).
#include "pamr_common.cuh"

namespace pamr {

int launch_propagate_tuned(const float* aff, const float* m_in, float* m_out, int B, int C, int H, int W,
                           const Dilations& dil, unsigned* cls_max, int dev, cudaStream_t s, bool* handled) {
    (void)aff; (void)m_in; (void)m_out; (void)B; (void)C; (void)H; (void)W; (void)dil; (void)cls_max; (void)dev; (void)s;
    *handled = false;
    return PAMR_OK;
}

}  // namespace pamr
