/*
 * pdlb200_pp.h — the reference-side half of the drop-in boundary, shared by
 *   (a) perl/PDL-B200/B200.xs        zero-touch attach: swaps these into live vtables, and
 *   (b) perl/PDL-B200/PP/Ops.pd      PDL::PP route: pp_def()s whose generated readdata/redodims
 *                                    bodies are one call into the functions below.
 * It copies what a generated readdata reads from pdl_trans (lib/PDL/Core/pdl.h.PL:381-403,471-482)
 * into the POD descriptor of include/pdlb200.h and calls pdlb200_readdata().
 *
 * Data store without touching Core: ndarray data lives in CUDA managed memory.  The redodims
 * hook gives every output the op creates a managed buffer BEFORE core's PDL_ENSURE_ALLOCATED
 * would allocate + zero-fill a Perl SV (pdlapi.c:14-17,172-209); large host-backed inputs are
 * moved into managed memory the first time a device op reads them (their Perl SV is released,
 * the pdl becomes PDL_DONTTOUCHDATA like an mmapped ndarray); small ones (<= STAGE_MAX bytes:
 * Perl scalars, 0-dim outputs stored inline in pdl.value) go through a pinned staging buffer.
 * Host code keeps dereferencing pdl->data; the driver migrates pages on demand (lazy host
 * sync) and chained ops never cross PCIe.
 *
 * Include after EXTERN.h/perl.h/XSUB.h, pdl.h, pdlcore.h.
 */
#ifndef PDLB200_PP_H
#define PDLB200_PP_H

#include "pdlb200.h"

#define PDLB200_STAGE_MAX 65536              /* parameters up to this many bytes are staged, not migrated */
#define PDLB200_STAGE_BYTES (8 * PDLB200_STAGE_MAX)

typedef pdl_error (*pdlb200_trans_fn)(pdl_trans *);

static int pdlb200_pp_enabled = 1;
static int pdlb200_pp_verbose = 0;
static unsigned long pdlb200_pp_device_calls = 0, pdlb200_pp_host_calls = 0, pdlb200_pp_migrated = 0, pdlb200_pp_staged = 0;
static char *pdlb200_pp_stage = NULL;        /* pinned, device-visible */
static size_t pdlb200_pp_stage_used = 0;

static int pdlb200_pp_init(void) {
  if (pdlb200_device_count() <= 0) return -1;
  if (!pdlb200_pp_stage) pdlb200_pp_stage = (char *)pdlb200_host_alloc(PDLB200_STAGE_BYTES);
  return pdlb200_pp_stage ? 0 : -2;
}

static void pdlb200_pp_free_managed(pdl *it, Size_t param) {
  (void)param;
  if (it->data) { pdlb200_managed_free(it->data); it->data = NULL; }
}

static uint64_t pdlb200_pp_badval_bits(Core *PDLc, pdl *p) {
  uint64_t bits = 0;
#define X(sym, ctype, ppsym, ...) \
  case sym: { ctype v = p->has_badvalue ? p->badvalue.value.ppsym : PDLc->bvals.ppsym; memcpy(&bits, &v, sizeof v <= 8 ? sizeof v : 8); } break;
  switch (p->datatype) {
    PDL_TYPELIST_REAL(X)
    default: break;
  }
#undef X
  return bits;
}
static int pdlb200_pp_badval_isnan(Core *PDLc, pdl *p) {
  if (p->datatype == PDL_F) { float v = p->has_badvalue ? p->badvalue.value.F : PDLc->bvals.F; return v != v; }
  if (p->datatype == PDL_D) { double v = p->has_badvalue ? p->badvalue.value.D : PDLc->bvals.D; return v != v; }
  return 0;
}

/* Move a physical ndarray's data into managed memory (once). */
static int pdlb200_pp_migrate(Core *PDLc, pdl *it) {
  dTHX;
  void *m;
  if (!it->data || it->nbytes <= 0) return 0;
  m = pdlb200_managed_alloc((size_t)it->nbytes);
  if (!m) return -1;
  memcpy(m, it->data, (size_t)it->nbytes);
  if (it->datasv) { SvREFCNT_dec((SV *)it->datasv); it->datasv = NULL; }
  it->data = m;
  it->state |= PDL_DONTTOUCHDATA | PDL_ALLOCATED;
  PDLc->add_deletedata_magic(it, pdlb200_pp_free_managed, 0);
  pdlb200_pp_migrated++;
  return 0;
}

typedef struct { pdl *owner; char *slot; size_t nbytes; } pdlb200_pp_staged_t;

/* Device-usable pointer for the buffer that physically holds p's data. */
static void *pdlb200_pp_device_view(Core *PDLc, pdl *p, int is_output, pdlb200_pp_staged_t *st, int *nst) {
  pdl *owner = PDL_VAFFOK(p) ? p->vafftrans->from : p;
  int kind;
  if (!owner->data) return NULL;
  kind = pdlb200_ptr_kind(owner->data);
  if (kind != 0) return owner->data;          /* managed, pinned or device memory already */
  if ((size_t)owner->nbytes <= PDLB200_STAGE_MAX && pdlb200_pp_stage_used + (size_t)owner->nbytes + 64 <= PDLB200_STAGE_BYTES) {
    int i;
    char *slot;
    for (i = 0; i < *nst; i++)
      if (st[i].owner == owner) {               /* aliasing parameters (inplace ops) share a slot */
        if (is_output) st[i].nbytes = (size_t)owner->nbytes;
        return st[i].slot;
      }
    slot = pdlb200_pp_stage + pdlb200_pp_stage_used;
    pdlb200_pp_stage_used += ((size_t)owner->nbytes + 63) & ~(size_t)63;
    memcpy(slot, owner->data, (size_t)owner->nbytes);
    st[*nst].owner = owner; st[*nst].slot = slot; st[*nst].nbytes = is_output ? (size_t)owner->nbytes : 0;
    (*nst)++;
    pdlb200_pp_staged++;
    return slot;
  }
  if (pdlb200_pp_migrate(PDLc, owner) != 0) return NULL;
  return owner->data;
}

/* redodims: the original (or the default) first, then managed buffers for the outputs this op creates. */
static pdl_error pdlb200_pp_redodims(Core *PDLc, pdl_trans *tr, pdlb200_trans_fn orig) {
  pdl_error PDL_err = {0, NULL, 0};
  PDL_Indx i;
  PDL_err = orig ? orig(tr) : PDLc->redodims_default(tr);
  if (PDL_err.error || !pdlb200_pp_enabled || tr->__datatype > PDL_D) return PDL_err;
  for (i = tr->vtable->nparents; i < tr->vtable->npdls; i++) {
    pdl *o = tr->pdls[i];
    PDL_Indx nbytes;
    if (!o || (o->state & PDL_ALLOCATED) || o->data || o->datatype > PDL_D || o->nvals <= 0) continue;
    if (o->trans_parent != tr) continue;       /* only ndarrays this op creates */
    nbytes = o->nvals * (PDL_Indx)PDLc->howbig(o->datatype);
    if (nbytes <= PDLB200_STAGE_MAX) continue;  /* small outputs keep core's inline / SV storage */
    {
      void *m = pdlb200_managed_alloc((size_t)nbytes);
      if (!m) continue;                        /* core will allocate host memory; we migrate later */
      o->data = m; o->nbytes = nbytes;
      o->state |= PDL_ALLOCATED | PDL_DONTTOUCHDATA;
      PDLc->add_deletedata_magic(o, pdlb200_pp_free_managed, 0);
    }
  }
  return PDL_err;
}

/* readdata: `fallback` (may be NULL) is the reference's own readdata for what is not on the device path. */
static pdl_error pdlb200_pp_readdata(Core *PDLc, pdl_trans *tr, int opid, pdlb200_trans_fn fallback) {
  pdl_error PDL_err = {0, NULL, 0};
  pdl_transvtable *vt = tr->vtable;
  pdlb200_trans d;
  pdlb200_pp_staged_t st[PDLB200_MAXPDLS];
  int nst = 0, rc, on_device = pdlb200_pp_enabled;
  int32_t anybad = 0;
  PDL_Indx i, j, npdls = vt->npdls;
  char err[512];
  if (tr->__datatype > PDL_D || npdls > PDLB200_MAXPDLS || tr->broadcast.ndims > PDLB200_MAXDIMS) on_device = 0;
  for (j = 0; on_device && j < npdls; j++) if (tr->pdls[j]->datatype > PDL_D) on_device = 0;
  if (!on_device) {
    if (fallback) { pdlb200_pp_host_calls++; return fallback(tr); }
    return PDLc->make_error(PDL_EUSERERROR, "PDL::B200 %s: type %d is outside the device type matrix and no host body is attached",
                            vt->name, (int)tr->__datatype);
  }
  if (pdlb200_pp_init() != 0)
    return PDLc->make_error(PDL_EFATAL, "PDL::B200 %s: no CUDA device available and there is no CPU fallback", vt->name);

  memset(&d, 0, sizeof d);
  d.op = opid; d.datatype = tr->__datatype; d.bvalflag = tr->bvalflag;
  d.npdls = (int32_t)npdls; d.ndims = (int32_t)tr->broadcast.ndims;
  for (i = 0; i < tr->broadcast.ndims; i++) {
    d.dims[i] = tr->broadcast.dims[i];
    for (j = 0; j < npdls; j++) d.incs[i * npdls + j] = PDL_BRC_INC(tr->broadcast.incs, npdls, j, i);
  }
  if (opid == PDLB200_OP_MATMULT) {
    /* ind_sizes are sorted by name: h, t, w ([gen] Primitive-pp-matmult.c); the ABI wants t, h, w */
    d.ind[0] = tr->ind_sizes[1]; d.ind[1] = tr->ind_sizes[0]; d.ind[2] = tr->ind_sizes[2];
    for (i = 0; i < 6; i++) d.rinc[i] = tr->inc_sizes[i];
  } else if (opid == PDLB200_OP_OUTER) {
    /* ind_names are sorted (pdl.h.PL:396): m, n; the ABI wants n, m.  inc_sizes follow the parameter order:
     * a(n), b(m), c(n,m) = exactly rinc[0..3] */
    d.ind[0] = tr->ind_sizes[1]; d.ind[1] = tr->ind_sizes[0];
    for (i = 0; i < 4; i++) d.rinc[i] = tr->inc_sizes[i];
  } else if (vt->ninds >= 1) {
    d.ind[0] = tr->ind_sizes[0];
    for (i = 0; i < vt->nind_ids && i < 8; i++) d.rinc[i] = tr->inc_sizes[i];
  }
  /* OtherPars: setvaltobad(double value) / setbadtoval(double newval) — the params struct holds that one double
   * ([gen] Bad-pp-setvaltobad.c: typedef struct pdl_params_setvaltobad { double value; }) */
  if ((opid == PDLB200_OP_SETVALTOBAD || opid == PDLB200_OP_SETBADTOVAL) && tr->params) d.param = *(double *)tr->params;
  d.anybad = &anybad;
  pdlb200_pp_stage_used = 0;
  for (j = 0; j < npdls; j++) {
    pdl *p = tr->pdls[j];
    int is_out = j >= vt->nparents;
    void *base = NULL;
    if (p->nvals > 0) {
      base = pdlb200_pp_device_view(PDLc, p, is_out, st, &nst);
      if (!base) return PDLc->make_error(PDL_EFATAL, "PDL::B200 %s: no device-usable storage for parameter %s", vt->name, vt->par_names[j]);
    }
    d.pdls[j].data = base;
    d.pdls[j].offs = PDL_REPROFFS(p);
    d.pdls[j].type = p->datatype;
    d.pdls[j].badval = pdlb200_pp_badval_bits(PDLc, p);
    d.pdls[j].flags = ((p->state & PDL_BADVAL) ? PDLB200_PAR_BADFLAG : 0) | (pdlb200_pp_badval_isnan(PDLc, p) ? PDLB200_PAR_BADNAN : 0);
  }
  rc = pdlb200_readdata(&d, err, sizeof err);
  /* no CPU fallback for the device type matrix: an unsupported shape (e.g. > 8 unmergeable broadcast dims) is an error */
  if (rc == 0) rc = pdlb200_sync(NULL, err, sizeof err);   /* host code may read pdl->data as soon as we return */
  if (rc != 0) return PDLc->make_error(PDL_EUSERERROR, "PDL::B200 %s: %s", vt->name, err);
  for (i = 0; i < nst; i++)
    if (st[i].nbytes) memcpy(st[i].owner->data, st[i].slot, st[i].nbytes);
  /* outputs flagged BAD by the op itself: minimum/maximum(_ind) with no good element (Ufunc.pd:463-464) */
  if (opid >= PDLB200_OP_MINIMUM && opid <= PDLB200_OP_MAXIMUM_IND && !tr->bvalflag && tr->ind_sizes[0] == 0)
    tr->pdls[1]->state |= PDL_BADVAL;
  /* $PDLSTATESETBAD / $PDLSTATESETGOOD of the Bad.pd bodies (Bad.pd:634,676,695-707,805,838,857) */
  if (opid == PDLB200_OP_SETBADIF || opid == PDLB200_OP_SETVALTOBAD ||
      (anybad && opid >= PDLB200_OP_SETNANTOBAD && opid <= PDLB200_OP_SETNONFINITETOBAD))
    tr->pdls[npdls - 1]->state |= PDL_BADVAL;
  else if (opid == PDLB200_OP_SETBADTONAN || opid == PDLB200_OP_SETBADTOVAL || opid == PDLB200_OP_BADMASK)
    tr->pdls[npdls - 1]->state &= ~PDL_BADVAL;
  /* minmaximum: a row without a usable element marks all four outputs BAD (Ufunc.pd:578-583) */
  if (opid == PDLB200_OP_MINMAXIMUM && anybad)
    for (j = vt->nparents; j < npdls; j++) tr->pdls[j]->state |= PDL_BADVAL;
  pdlb200_pp_device_calls++;
  if (pdlb200_pp_verbose) fprintf(stderr, "PDL::B200 %s -> %s\n", vt->name, pdlb200_last_kernel());
  return PDL_err;
}

#endif /* PDLB200_PP_H */
