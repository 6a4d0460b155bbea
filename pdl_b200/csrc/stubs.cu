// stubs.cu — launchers not yet on the device path report so loudly (no CPU fallback).
#include "common.cuh"
namespace pdlb200 {
int launch_scan(const pdlb200_trans *t, const Err &E) {
  return E.fail(PDLB200_EUNSUPPORTED, "%s: scans are not on the device path yet", pdlb200_op_name(t->op));
}
}  // namespace pdlb200
