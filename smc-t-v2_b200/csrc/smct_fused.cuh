// smct_fused.cuh — fused persistent forward (placeholder until the fused kernel lands).
#pragma once
#include "smct_common.cuh"
#include "../../include/smct.h"
namespace smct {
inline size_t fused_forward_ws(const smct_shape&) { return 0; }
inline bool fused_supported(const smct_shape&, const smct_io&) { return false; }
inline int forward_fused(const smct_shape&, const smct_params&, const smct_io&, void*, size_t, cudaStream_t, const char** err) {
  *err = "fused forward not built"; return SMCT_ERR_UNSUPPORTED;
}
}  // namespace smct
