// tcgen05 / TMEM deformable-convolution forward (bf16 and bf16x3 operand splits).  Placeholder until the
// tensor-core kernel lands: reports D2B_EUNSUPPORTED so that callers asking for precision != 0 fail loudly
// instead of silently running a different code path.
#include "common.cuh"

int d2b_deform_conv_forward_tc(const float*, const float*, const float*, const float*, const float*,
                               const d2b_dcn_params*, int, float*, void*) {
  return D2B_EUNSUPPORTED;
}
