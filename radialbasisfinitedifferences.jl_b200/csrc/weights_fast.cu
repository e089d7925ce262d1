// weights_fast.cu -- register/DMMA fast path of the weight kernel (collocated rows, one RHS set per
// factorisation).  Placeholder until the tuned kernel lands: reports "unsupported" so the driver in
// weights.cu uses the generic shared-memory kernel.
#include "common.cuh"
#include "tables.cuh"

int rbffd_weights_fast(rbffd_context*, const OpTables&, const double*, int64_t, const double*, int64_t,
                       const int32_t*, int32_t*, double*, int*) {
    return RBFFD_ERR_UNSUPPORTED;
}
