// reduce_plan.cu — host planner for the Ufunc reductions: collapse the broadcast
// dims, pick the cooperation width (thread / warp / CTA per row) and, when rows are
// too few to fill the GPU, the number of chunks each row is cut into.
#include <cstdlib>
#include <cstring>
#include "reduce.cuh"

namespace pdlb200 {

int rd_build_plan(const pdlb200_trans *t, size_t in_size, size_t out_size, size_t acc_size,
                  RdPlan *p, RdLaunch *l, const Err &E, bool heavy_row_end, int blocks_per_sm) {
  if (t->npdls != 2)
    return E.fail(PDLB200_EINVAL, "%s: expected 2 parameters, got %d", pdlb200_op_name(t->op), t->npdls);
  if (t->ind[0] < 0) return E.fail(PDLB200_EINVAL, "%s: size of dim n is %lld", pdlb200_op_name(t->op), (long long)t->ind[0]);
  Collapsed c;
  collapse_dims(t, &c);
  if (c.nd > MAXD)
    return E.fail(PDLB200_EUNSUPPORTED, "%s: %d non-mergeable broadcast dims exceed the device walker's %d",
                  pdlb200_op_name(t->op), c.nd, MAXD);
  memset(p, 0, sizeof *p);
  p->n = t->ind[0];
  p->inc_n = t->rinc[0];
  p->goff = t->ind[1]; p->inc_r = t->rinc[1];   // read by the PART_* reducers only (reduce_shard.cu)
  p->nrows = c.total;
  p->nd = c.nd;
  for (int d = 0; d < c.nd; d++) { p->dims[d] = c.dims[d]; p->sa[d] = c.st[0][d]; p->sb[d] = c.st[1][d]; }
  if (c.total > 0 && (!t->pdls[1].data || (p->n > 0 && !t->pdls[0].data)))
    return E.fail(PDLB200_EINVAL, "%s: parameter got NULL data", pdlb200_op_name(t->op));
  p->a = (const char *)t->pdls[0].data + t->pdls[0].offs * (int64_t)in_size;
  p->b = (char *)t->pdls[1].data + t->pdls[1].offs * (int64_t)out_size;
  p->abad = t->pdls[0].badval; p->bbad = t->pdls[1].badval;
  p->abadnan = (t->pdls[0].flags & PDLB200_PAR_BADNAN) != 0;
  p->badmode = t->bvalflag != 0;
  p->chunk = p->n; p->nchunks = 1;
  if (p->nrows == 0) return PDLB200_OK;

  const int sms = sm_count();
  const int64_t n = p->n;
  // "column" reduction: the reduced dim is strided but broadcast dim 0 is unit-stride in a,
  // so one thread per row makes every warp-wide load a contiguous run.
  const bool column = (p->inc_n != 1 || n <= 1) && (p->sa[0] == 1 || p->sa[0] == -1) && p->dims[0] >= 32;
  int mode;
  if (column) mode = 0;
  else if ((int64_t)n * (int64_t)in_size >= 32768) {
    mode = 2;
    if (heavy_row_end) {
      // min/max reducers end every row with a tree over (value, index, state) records and a re-walk test: a CTA needs
      // several trips per thread to amortise it (measured, float rows: 32 KB rows 0.81 CTA-per-row vs 1.02
      // warp-per-row; 64 KB rows of cfg2 0.98 vs 1.03).  Warp-per-row only while every resident warp gets (almost)
      // the same number of rows — otherwise the last wave costs more than the row ends save.
      const int64_t row_bytes = (int64_t)n * (int64_t)in_size;
      const double r = (double)p->nrows / (double)((int64_t)sms * 8 * (RD_THREADS / 32));   // rows per resident warp
      if (row_bytes < 65536) { if (r >= 0.5) mode = 1; }
      else if (row_bytes <= 262144) { if (r >= 1.0 && (double)(int64_t)(r + 0.999999) / r <= 1.03) mode = 1; }
    }
  }
  else if (n >= 64) mode = 1;
  else mode = 0;
  if (const char *e = getenv("PDLB200_REDUCE_MODE")) { int m = atoi(e); if (m >= 0 && m <= 2) mode = m; }

  const int64_t rows_per_cta = mode == 0 ? RD_THREADS : mode == 1 ? RD_THREADS / 32 : 1;
  int64_t ctas = (p->nrows + rows_per_cta - 1) / rows_per_cta;
  // the CTAs that are resident at once (launch bounds: 8 per SM for the light reducers, 5 for the heavy ones)
  const int64_t target = (int64_t)sms * (blocks_per_sm > 0 ? blocks_per_sm : 8);
  // too few row-CTAs: cut n into chunks (>= 16K elements each so the partial traffic stays negligible).  The grid must
  // not spill into a second, nearly empty wave — 64 rows x 19 chunks = 1216 CTAs on 1184 slots ran at 0.80 of peak,
  // the last 32 CTAs alone on the machine — so the chunk count is rounded DOWN to what is resident at once.
  const int64_t min_chunk = mode == 0 ? 256 : 16384;
  if (ctas < target / 2 && n >= 2 * min_chunk) {
    int64_t want = target / ctas;
    int64_t maxc = n / min_chunk;
    if (want > maxc) want = maxc;
    if (want > 65535) want = 65535;
    if (want > 1) {
      int64_t chunk = (n + want - 1) / want;
      chunk = (chunk + 4095) / 4096 * 4096;   // keeps 16-byte alignment of chunk starts for every type
      p->chunk = chunk;
      p->nchunks = (int)((n + chunk - 1) / chunk);
    }
  }
  // per-thread element indices are 32-bit relative to the chunk start: keep chunks below 2^30
  if (p->chunk > (1ll << 30)) {
    const int64_t want = (n + (1ll << 30) - 1) >> 30;
    int64_t chunk = (n + want - 1) / want; chunk = (chunk + 4095) / 4096 * 4096;
    p->chunk = chunk; p->nchunks = (int)((n + chunk - 1) / chunk);
  }
  if (const char *e = getenv("PDLB200_REDUCE_CHUNKS")) {
    int want = atoi(e);
    if (want >= 1 && want <= 65535 && n > 0) {
      int64_t chunk = (n + want - 1) / want; chunk = (chunk + 4095) / 4096 * 4096;
      p->chunk = chunk; p->nchunks = (int)((n + chunk - 1) / chunk);
    }
  }
  if (p->nchunks > 1) {
    p->partial = (char *)scratch((size_t)p->nrows * p->nchunks * acc_size, (cudaStream_t)t->stream);
    if (!p->partial) return E.fail(PDLB200_ECUDA, "%s: cannot allocate %zu bytes of reduction scratch",
                                   pdlb200_op_name(t->op), (size_t)p->nrows * p->nchunks * acc_size);
  }
  int64_t gx = ctas;
  // whole rows: up to 8 CTAs per SM in the grid even when fewer are resident (the queued ones even out the end of the
  // launch); rows cut into chunks: exactly what is resident
  const int64_t cap = p->nchunks > 1 ? target / p->nchunks : (int64_t)sms * 8;
  if (gx > cap) gx = cap < 1 ? 1 : cap;
  l->mode = mode;
  l->grid = dim3((unsigned)gx, (unsigned)p->nchunks, 1);
  return PDLB200_OK;
}

}  // namespace pdlb200
