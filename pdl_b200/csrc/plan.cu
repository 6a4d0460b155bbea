// plan.cu — host-side planning shared by all kernel families: collapse the
// broadcast dims of a descriptor and build the elementwise kernel's plan.
// This is the device-side replacement for what PDL_BROADCASTLOOP_START reads out of
// pdl_broadcast at run time (lib/PDL/Core/pdl.h.PL:640-667): dims, per-pdl incs, offsets.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include "elementwise.cuh"

namespace pdlb200 {

int Err::fail(int code, const char *fmt, ...) const {
  if (buf && len) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(buf, len, fmt, ap);
    va_end(ap);
  }
  return code;
}

void collapse_dims(const pdlb200_trans *t, Collapsed *c) {
  const int np = t->npdls;
  c->nd = 0; c->total = 1;
  for (int d = 0; d < t->ndims; d++) c->total *= t->dims[d];
  for (int d = 0; d < t->ndims; d++) {
    const int64_t n = t->dims[d];
    if (n == 1) continue;  // size-1 dims contribute nothing (their incs are ignored)
    if (c->nd > 0) {
      const int k = c->nd - 1;
      bool merge = true;
      for (int p = 0; p < np; p++)
        if (t->incs[d * np + p] != c->st[p][k] * c->dims[k]) { merge = false; break; }
      if (merge) { c->dims[k] *= n; continue; }
    }
    c->dims[c->nd] = n;
    for (int p = 0; p < np; p++) c->st[p][c->nd] = t->incs[d * np + p];
    c->nd++;
  }
  if (c->nd == 0) {
    c->nd = 1; c->dims[0] = 1;
    for (int p = 0; p < np; p++) c->st[p][0] = 0;
  }
}

int ew_build_plan(const pdlb200_trans *t, int nin, size_t in_size, size_t out_size,
                  bool state_checked_bad, EwPlan *p, const Err &E, size_t b_size) {
  if (b_size == 0) b_size = in_size;
  if (t->npdls != nin + 1)
    return E.fail(PDLB200_EINVAL, "%s: expected %d parameters, got %d", pdlb200_op_name(t->op), nin + 1, t->npdls);
  Collapsed c;
  collapse_dims(t, &c);
  if (c.nd > MAXD)
    return E.fail(PDLB200_EUNSUPPORTED, "%s: %d non-mergeable broadcast dims exceed the device walker's %d",
                  pdlb200_op_name(t->op), c.nd, MAXD);
  memset(p, 0, sizeof *p);
  size_t wide = in_size > out_size ? in_size : out_size;
  if (nin > 1 && b_size > wide) wide = b_size;
  const int VEC = (int)(16 / wide);
  p->nd = c.nd;
  for (int d = 0; d < c.nd; d++) p->dims[d] = c.dims[d];
  for (int k = 0; k <= nin; k++) {
    const pdlb200_par &par = t->pdls[k];
    const size_t sz = k == nin ? out_size : (k == 1 ? b_size : in_size);
    if (c.total > 0 && !par.data)
      return E.fail(PDLB200_EINVAL, "%s: parameter %d got NULL data", pdlb200_op_name(t->op), k);
    p->ptr[k] = (char *)par.data + par.offs * (int64_t)sz;
    for (int d = 0; d < c.nd; d++) p->st[k][d] = c.st[k][d];
    p->bad[k] = par.badval;
    p->badnan[k] = (par.flags & PDLB200_PAR_BADNAN) != 0;
    p->badchk[k] = state_checked_bad ? ((par.flags & PDLB200_PAR_BADFLAG) != 0) : 1;
    // vector access: unit stride along dim 0, base and every outer stride aligned to the access width
    const size_t bytes = (size_t)VEC * sz;
    bool ok = (c.st[k][0] == 1) && (((uintptr_t)p->ptr[k]) % bytes == 0);
    for (int d = 1; d < c.nd && ok; d++)
      if (((c.st[k][d] * (int64_t)sz) % (int64_t)bytes) != 0) ok = false;
    p->vec[k] = ok;
  }
  if (c.total == 0) { p->n_units = 0; return PDLB200_OK; }
  p->vpr = (c.dims[0] + VEC - 1) / VEC;
  p->n_units = p->vpr * (c.total / c.dims[0]);
  return PDLB200_OK;
}

int ew_grid(int64_t n_units, int unroll, const void *kernel) {
  // Persistent-style sizing: resident CTAs per SM (occupancy API) x SM count, capped by the work.
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, EW_THREADS, 0) != cudaSuccess || per_sm < 1)
    per_sm = 4;
  const int64_t full = (int64_t)sm_count() * per_sm;
  const int64_t need = (n_units + (int64_t)EW_THREADS * unroll - 1) / ((int64_t)EW_THREADS * unroll);
  int64_t g = need < full ? need : full;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace pdlb200
