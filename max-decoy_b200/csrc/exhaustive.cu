// exhaustive.cu -- MD_DECOY_EXHAUSTIVE (placeholder until the enumeration kernel lands)
#include "cubx.cuh"
void decoys_exhaustive_dev(md_ctx* ctx, uint32_t n, uint32_t n_per) {
  (void)ctx; (void)n; (void)n_per;
  throw MdError{MD_ERR_UNSUPPORTED, "MD_DECOY_EXHAUSTIVE is not implemented on the GPU yet"};
}
