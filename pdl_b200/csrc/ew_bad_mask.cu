// ew_bad_mask.cu — badmask, copybad (lib/PDL/Bad.pd:842-905, all types) and the floating-point-only
// setnantobad / setinftobad / setnonfinitetobad / setbadtonan (:679-806).
#include "ew_badops.cuh"
namespace pdlb200 {

// set{nan,inf,nonfinite}tobad: the output badflag is data-dependent (`if (flag) $PDLSTATESETBAD(b)`), so the
// call ends with a 4-byte read-back of a device flag word and a stream synchronise.
template <class Op, class T>
static int flagged(const pdlb200_trans *t, const char *name, const Err &E) {
  if (!t->anybad) return E.fail(PDLB200_EINVAL, "%s: pdlb200_trans.anybad must point to an int32", pdlb200_op_name(t->op));
  cudaStream_t s = (cudaStream_t)t->stream;
  int *flag = (int *)scratch(sizeof(int), s);
  if (!flag) return E.fail(PDLB200_ECUDA, "%s: cannot allocate the flag word", pdlb200_op_name(t->op));
  PDLB200_CUDA_OK(cudaMemsetAsync(flag, 0, sizeof(int), s), E);
  if (int rc = ew_launch_typed<Op, T, T, 1>(t, false, name, E, 0, flag)) return rc;
  if (t->tflags & PDLB200_TRANS_DEFER_ANYBAD) {     // pinned destination: no host round trip inside the call
    PDLB200_CUDA_OK(cudaMemcpyAsync(t->anybad, flag, sizeof(int), cudaMemcpyDeviceToHost, s), E);
    return PDLB200_OK;
  }
  int host = 0;
  PDLB200_CUDA_OK(cudaMemcpyAsync(&host, flag, sizeof(int), cudaMemcpyDeviceToHost, s), E);
  PDLB200_CUDA_OK(cudaStreamSynchronize(s), E);
  *t->anybad = host != 0;
  return PDLB200_OK;
}

int ew_bad_mask(const pdlb200_trans *t, const Err &E) {
  switch (t->op) {
#define Q(T) return ew_launch_typed<OpBadmask, T, T, 2>(t, false, "ew_badmask", E);
    case PDLB200_OP_BADMASK: switch (t->datatype) { PDLB200_BAD_CASES(Q) default: break; } break;
#undef Q
#define Q(T) return ew_launch_typed<OpCopybad, T, T, 2>(t, false, "ew_copybad", E);
    case PDLB200_OP_COPYBAD: switch (t->datatype) { PDLB200_BAD_CASES(Q) default: break; } break;
#undef Q
    case PDLB200_OP_SETNANTOBAD:
      if (t->datatype == PDLB200_F) return flagged<OpSetnantobad, float>(t, "ew_setnantobad", E);
      if (t->datatype == PDLB200_D) return flagged<OpSetnantobad, double>(t, "ew_setnantobad", E);
      break;
    case PDLB200_OP_SETINFTOBAD:
      if (t->datatype == PDLB200_F) return flagged<OpSetinftobad, float>(t, "ew_setinftobad", E);
      if (t->datatype == PDLB200_D) return flagged<OpSetinftobad, double>(t, "ew_setinftobad", E);
      break;
    case PDLB200_OP_SETNONFINITETOBAD:
      if (t->datatype == PDLB200_F) return flagged<OpSetnonfinitetobad, float>(t, "ew_setnonfinitetobad", E);
      if (t->datatype == PDLB200_D) return flagged<OpSetnonfinitetobad, double>(t, "ew_setnonfinitetobad", E);
      break;
    case PDLB200_OP_SETBADTONAN: {
      pdlb200_trans u = *t;
      u.bvalflag = 1;   // $ISBAD(a()) is compiled into both code copies (Bad.pd:795)
      if (t->datatype == PDLB200_F) return ew_launch_typed<OpSetbadtonan, float, float, 1>(&u, false, "ew_setbadtonan", E);
      if (t->datatype == PDLB200_D) return ew_launch_typed<OpSetbadtonan, double, double, 1>(&u, false, "ew_setbadtonan", E);
      break;
    }
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not on the device path", pdlb200_op_name(t->op), t->datatype);
}

int ew_bad_query(const pdlb200_trans *, const Err &);
int ew_bad_set(const pdlb200_trans *, const Err &);
int launch_badops(const pdlb200_trans *t, const Err &E) {
  if (t->op <= PDLB200_OP_ISNAN) return ew_bad_query(t, E);
  if (t->op == PDLB200_OP_SETBADIF || t->op == PDLB200_OP_SETVALTOBAD || t->op == PDLB200_OP_SETBADTOVAL) return ew_bad_set(t, E);
  return ew_bad_mask(t, E);
}
}  // namespace pdlb200
