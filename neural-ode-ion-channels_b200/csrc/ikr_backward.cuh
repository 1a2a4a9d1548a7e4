// ikr_backward.cuh -- backward sweep (placeholder until the adjoint kernel lands)
#ifndef IKR_BACKWARD_CUH_
#define IKR_BACKWARD_CUH_
#include "../../include/ikr.h"
#include "ikr_device.cuh"
namespace ikr {
inline size_t bwd_workspace_bytes(const ikr_desc*, long long) { return 0; }
inline int bwd_dispatch(const ikr_desc*, const ikr_io*, const ikr_bwd_io*, void*, size_t,
                        cudaStream_t) { return IKR_ERR_UNSUPPORTED; }
}  // namespace ikr
#endif
