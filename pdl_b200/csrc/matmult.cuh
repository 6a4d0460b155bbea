// matmult.cuh — plan shared by the two matmult kernels (matmult.cu, matmult_dmma.cu).
#pragma once
#include "common.cuh"

namespace pdlb200 {

constexpr int64_t MM_MAXZ = 65535;  // gridDim.z limit

struct MmPlan {
  const char *a, *b; char *c;      // bases with offs applied
  int64_t T, H, W;                 // sizes of t, h, w
  int64_t iat, iah, ibw, ibt, icw, ich;  // element strides
  int64_t dims[MAXD];              // collapsed broadcast (batch) dims
  int64_t sa[MAXD], sb[MAXD], sc[MAXD];
  int64_t nbatch;
  int64_t z0;                      // first broadcast position of this launch (grids are cut at 65535 in z)
  uint64_t abad, bbad, cbad;
  int nd;
  int abadnan, bbadnan, cbadnan;
};

int launch_matmult_dmma(const pdlb200_trans *t, const MmPlan &p, const Err &E);  // matmult_dmma.cu
int launch_matmult_tma(const pdlb200_trans *t, const MmPlan &p, const Err &E);   // matmult_tma.cu

}  // namespace pdlb200
