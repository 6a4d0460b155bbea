/*
 * pdlb200_pp.h — the reference-side half of the drop-in boundary, shared by
 *   (a) perl/PDL-B200/B200.xs        zero-touch attach: swaps these into live vtables, and
 *   (b) perl/PDL-B200/PP/Ops.pd      PDL::PP route: pp_def()s whose generated readdata/redodims
 *                                    bodies are one call into the functions below.
 * It copies what a generated readdata reads from pdl_trans (lib/PDL/Core/pdl.h.PL:381-403,471-482)
 * into the POD descriptor of include/pdlb200.h and calls pdlb200_readdata().
 *
 * Data store (north-star subsystem 1) without touching Core: the library's coherent store
 * (pdlb200_mbuf_*, include/pdlb200.h).  An ndarray the device path creates or adopts has a cudaMalloc'd
 * device buffer + a protected host mirror that IS pdl->data + dirty bits:
 *   - redodims hook: every output the op creates gets a store buffer BEFORE core's PDL_ENSURE_ALLOCATED
 *     would allocate + zero-fill a Perl SV (pdlapi.c:14-17,172-209);
 *   - readdata hook: parameters already in the store hand out their device pointer (a host-modified one is
 *     uploaded first); large host-backed ones are ADOPTED on first use (one H2D copy, the SV is released,
 *     the pdl becomes PDL_DONTTOUCHDATA like an mmapped ndarray); small ones (<= stage_max bytes: Perl
 *     scalars, 0-dim outputs stored inline in pdl.value) go through a pinned staging ring; ndarrays whose
 *     memory belongs to someone else (mmapped files, shared SVs) are staged through a temporary device
 *     buffer and written back;
 *   - NO synchronisation after the launch unless a result had to be copied back to plain host memory:
 *     chained ops on device-resident ndarrays are back-to-back launches;
 *   - host access: CPU transformations are seen in the Core function table (PDL->make_trans_mutual) and
 *     their parameters made current first; every other dereference of pdl->data (lib/PDL/Core.xs at /
 *     listref / sclr / stringify, pdlconv.c vaffine read/writeback) faults into the store, which
 *     downloads the buffer once and resumes.
 *
 * Include after EXTERN.h/perl.h/XSUB.h, pdl.h, pdlcore.h.
 */
#ifndef PDLB200_PP_H
#define PDLB200_PP_H

#include <complex.h>
#include "pdlb200.h"

#define PDLB200_STAGE_DEFAULT 65536          /* parameters up to this many bytes are staged, not adopted */
#define PDLB200_STAGE_BYTES (1 << 20)        /* pinned staging ring */

typedef pdl_error (*pdlb200_trans_fn)(pdl_trans *);

static int pdlb200_pp_enabled = 1;
static int pdlb200_pp_verbose = 0;
static unsigned long pdlb200_pp_device_calls = 0, pdlb200_pp_host_calls = 0, pdlb200_pp_migrated = 0, pdlb200_pp_staged = 0,
                     pdlb200_pp_syncs = 0, pdlb200_pp_transient = 0, pdlb200_pp_cpu_trans = 0;
static char *pdlb200_pp_stage = NULL;        /* pinned, device-visible */
static size_t pdlb200_pp_stage_used = 0;
static size_t pdlb200_pp_stage_max = PDLB200_STAGE_DEFAULT;

static int pdlb200_pp_init(void) {
  if (pdlb200_device_count() <= 0) return -1;
  if (!pdlb200_pp_stage) {
    const char *e = getenv("PDLB200_STAGE_MAX");
    if (e) { long v = atol(e); if (v >= 0 && v <= PDLB200_STAGE_BYTES / 8) pdlb200_pp_stage_max = (size_t)v; }
    pdlb200_pp_stage = (char *)pdlb200_host_alloc(PDLB200_STAGE_BYTES);
  }
  return pdlb200_pp_stage ? 0 : -2;
}

/* delete-data magic of every ndarray whose data lives in the store; a pdl that was handed back to plain host
 * memory (pdlb200_pp_release) keeps the magic, its data pointer is then no store buffer any more */
static void pdlb200_pp_free_store(pdl *it, Size_t param) {
  (void)param;
  if (it->data && pdlb200_mbuf_is(it->data)) { pdlb200_mbuf_free(it->data); it->data = NULL; }
}

/* complex float / complex double are on the device path for plus minus mult divide (include/pdlb200.h) */
static int pdlb200_pp_type_ok(int datatype, int opid) {
  if (datatype <= PDL_D) return 1;
  return (datatype == PDL_CF || datatype == PDL_CD) && opid >= PDLB200_OP_PLUS && opid <= PDLB200_OP_DIVIDE;
}

static uint64_t pdlb200_pp_badval_bits(Core *PDLc, pdl *p) {
  uint64_t bits = 0;
  if (p->datatype == PDL_CF) {            /* both parts, real part in the low 4 bytes */
    complex float v = p->has_badvalue ? p->badvalue.value.G : PDLc->bvals.G;
    memcpy(&bits, &v, 8);
    return bits;
  }
  if (p->datatype == PDL_CD) {            /* the real part's bits (the ABI needs equal parts: checked by the caller) */
    complex double v = p->has_badvalue ? p->badvalue.value.C : PDLc->bvals.C;
    double re = creal(v);
    memcpy(&bits, &re, 8);
    return bits;
  }
#define X(sym, ctype, ppsym, ...) \
  case sym: { ctype v = p->has_badvalue ? p->badvalue.value.ppsym : PDLc->bvals.ppsym; memcpy(&bits, &v, sizeof v <= 8 ? sizeof v : 8); } break;
  switch (p->datatype) {
    PDL_TYPELIST_REAL(X)
    default: break;
  }
#undef X
  return bits;
}
static int pdlb200_pp_cd_badval_ok(Core *PDLc, pdl *p) {
  complex double v;
  if (p->datatype != PDL_CD) return 1;
  v = p->has_badvalue ? p->badvalue.value.C : PDLc->bvals.C;
  return creal(v) == cimag(v) || (creal(v) != creal(v) && cimag(v) != cimag(v));
}
static int pdlb200_pp_badval_isnan(Core *PDLc, pdl *p) {
  if (p->datatype == PDL_CF) { complex float v = p->has_badvalue ? p->badvalue.value.G : PDLc->bvals.G; return crealf(v) != crealf(v) || cimagf(v) != cimagf(v); }
  if (p->datatype == PDL_CD) { complex double v = p->has_badvalue ? p->badvalue.value.C : PDLc->bvals.C; return creal(v) != creal(v) || cimag(v) != cimag(v); }
  if (p->datatype == PDL_F) { float v = p->has_badvalue ? p->badvalue.value.F : PDLc->bvals.F; return v != v; }
  if (p->datatype == PDL_D) { double v = p->has_badvalue ? p->badvalue.value.D : PDLc->bvals.D; return v != v; }
  return 0;
}

/* May this physical ndarray's data be re-homed into the store?  Not if someone else owns the memory: an
 * mmapped file or shared-memory ndarray (PDL_DONTTOUCHDATA set by its creator, lib/PDL/Core.xs:1120-1143,
 * set_data_by_file_map) or an SV that Perl code holds a reference to (get_dataref, Core.xs:1145-1160) —
 * writes must keep landing in THAT memory. */
static int pdlb200_pp_adoptable(pdl *it) {
  dTHX;
  if (it->state & PDL_DONTTOUCHDATA) return 0;
  if (!it->datasv) return 0;
  if (SvREFCNT((SV *)it->datasv) != 1) return 0;
  if (PDL_ISMAGIC(it)) return 0;                /* pthread / delete-data magic from someone else */
  return 1;
}

/* Re-home a physical ndarray's data into the store (once): one H2D copy, the SV is released. */
static int pdlb200_pp_adopt(Core *PDLc, pdl *it, int upload, char *err, size_t errlen) {
  dTHX;
  void *m = upload ? pdlb200_mbuf_adopt(it->data, (size_t)it->nbytes, err, errlen) : pdlb200_mbuf_new((size_t)it->nbytes);
  if (!m) return -1;
  SvREFCNT_dec((SV *)it->datasv); it->datasv = NULL;
  it->data = m;
  it->state |= PDL_DONTTOUCHDATA | PDL_ALLOCATED;
  PDLc->add_deletedata_magic(it, pdlb200_pp_free_store, 0);
  pdlb200_pp_migrated++;
  return 0;
}

/* Hand an ndarray back to plain host memory (a fresh Perl SV): what get_dataref / setdims / reshape need,
 * which refuse PDL_DONTTOUCHDATA ndarrays (Core.xs:1147, pdlapi.c:183). */
static int pdlb200_pp_release(Core *PDLc, pdl *it) {
  dTHX;
  char err[256];
  SV *sv;
  void *m = it->data;
  (void)PDLc;
  if (!m || !pdlb200_mbuf_is(m)) return 0;
  if (pdlb200_mbuf_host(m, 0, err, sizeof err) != 0) return -1;
  if ((size_t)it->nbytes > sizeof(it->value)) {
    sv = newSVpvn("", 0);
    (void)SvGROW(sv, (STRLEN)it->nbytes + 1);
    SvCUR_set(sv, (STRLEN)it->nbytes);
    memcpy(SvPVX(sv), m, (size_t)it->nbytes);
    it->datasv = sv; it->data = SvPVX(sv);
  } else {
    memcpy(&it->value, m, (size_t)it->nbytes);
    it->data = &it->value;
  }
  it->state &= ~PDL_DONTTOUCHDATA;
  pdlb200_mbuf_free(m);
  return 0;
}

typedef struct { pdl *owner; char *slot; void *tmp; size_t nbytes; int writeback; } pdlb200_pp_staged_t;

/* Device-usable pointer for the buffer that physically holds p's data.
 * reads: the kernel reads it (an input, or an output it only partly overwrites); writes: the kernel writes it. */
static void *pdlb200_pp_device_view(Core *PDLc, pdl *p, int reads, int writes, pdlb200_pp_staged_t *st, int *nst,
                                    int *need_sync, char *err, size_t errlen) {
  pdl *owner = PDL_VAFFOK(p) ? p->vafftrans->from : p;
  size_t nbytes = (size_t)owner->nbytes;
  int i;
  if (!owner->data) { snprintf(err, errlen, "ndarray has no data"); return NULL; }
  if (pdlb200_mbuf_is(owner->data))
    return pdlb200_mbuf_dev(owner->data, writes, writes && !reads, NULL, err, errlen);
  for (i = 0; i < *nst; i++)
    if (st[i].owner == owner) {                 /* aliasing parameters (inplace ops) share a slot */
      if (writes) { st[i].writeback = 1; *need_sync = 1; }
      return st[i].tmp ? pdlb200_mbuf_dev(st[i].tmp, writes, 0, NULL, err, errlen) : (void *)st[i].slot;
    }
  if (nbytes <= pdlb200_pp_stage_max) {
    char *slot;
    size_t need = (nbytes + 63) & ~(size_t)63;
    if (pdlb200_pp_stage_used + need > PDLB200_STAGE_BYTES) {   /* ring wrap: earlier slots may still be read by kernels in flight */
      if (pdlb200_sync(NULL, err, errlen) != 0) return NULL;
      pdlb200_pp_syncs++;
      pdlb200_pp_stage_used = 0;                /* this call's earlier slots sit at the END of the ring: untouched by the new ones */
    }
    slot = pdlb200_pp_stage + pdlb200_pp_stage_used;
    pdlb200_pp_stage_used += need;
    if (reads) memcpy(slot, owner->data, nbytes);
    st[*nst].owner = owner; st[*nst].slot = slot; st[*nst].tmp = NULL; st[*nst].nbytes = nbytes; st[*nst].writeback = writes;
    (*nst)++;
    if (writes) *need_sync = 1;
    pdlb200_pp_staged++;
    return slot;
  }
  if (pdlb200_pp_adoptable(owner)) {
    if (pdlb200_pp_adopt(PDLc, owner, reads, err, errlen) != 0) return NULL;
    return pdlb200_mbuf_dev(owner->data, writes, writes && !reads, NULL, err, errlen);
  }
  {
    /* someone else's memory (mmapped file, shared SV): a temporary store buffer for this one call */
    void *tmp = reads ? pdlb200_mbuf_adopt(owner->data, nbytes, err, errlen) : pdlb200_mbuf_new(nbytes);
    if (!tmp) { if (!reads) snprintf(err, errlen, "cannot allocate %zu bytes on the device", nbytes); return NULL; }
    st[*nst].owner = owner; st[*nst].slot = NULL; st[*nst].tmp = tmp; st[*nst].nbytes = nbytes; st[*nst].writeback = writes;
    (*nst)++;
    if (writes) *need_sync = 1;
    pdlb200_pp_transient++;
    return pdlb200_mbuf_dev(tmp, writes, writes && !reads, NULL, err, errlen);
  }
}

/* after the launch: results that have to live in plain host memory are copied back (the only synchronisation) */
static int pdlb200_pp_finish_staged(pdlb200_pp_staged_t *st, int nst, int need_sync, char *err, size_t errlen) {
  int i, rc = 0;
  if (need_sync) { rc = pdlb200_sync(NULL, err, errlen); pdlb200_pp_syncs++; }
  for (i = 0; i < nst; i++) {
    if (st[i].tmp) {
      if (rc == 0 && st[i].writeback) {
        rc = pdlb200_mbuf_host(st[i].tmp, 0, err, errlen);
        if (rc == 0) memcpy(st[i].owner->data, st[i].tmp, st[i].nbytes);
      }
      pdlb200_mbuf_free(st[i].tmp);
    } else if (rc == 0 && st[i].writeback) memcpy(st[i].owner->data, st[i].slot, st[i].nbytes);
  }
  return rc;
}

/* store buffers for the outputs this transformation creates, BEFORE core would allocate + zero-fill an SV */
static void pdlb200_pp_give_outputs(Core *PDLc, pdl_trans *tr) {
  PDL_Indx i;
  if (!pdlb200_pp_enabled || tr->__datatype > PDL_CD || tr->__datatype == PDL_LD || pdlb200_pp_init() != 0) return;
  for (i = tr->vtable->nparents; i < tr->vtable->npdls; i++) {
    pdl *o = tr->pdls[i];
    PDL_Indx nbytes;
    if (!o || (o->state & PDL_ALLOCATED) || o->data || o->datatype > PDL_CD || o->datatype == PDL_LD || o->nvals <= 0) continue;
    if (o->trans_parent != tr) continue;       /* only ndarrays this op creates */
    nbytes = o->nvals * (PDL_Indx)PDLc->howbig(o->datatype);
    if ((size_t)nbytes <= pdlb200_pp_stage_max || (size_t)nbytes <= sizeof(o->value)) continue;  /* small outputs keep core's inline / SV storage */
    {
      void *m = pdlb200_mbuf_new((size_t)nbytes);
      if (!m) continue;                        /* core will allocate host memory; adopted later */
      o->data = m; o->nbytes = nbytes;
      o->state |= PDL_ALLOCATED | PDL_DONTTOUCHDATA;
      PDLc->add_deletedata_magic(o, pdlb200_pp_free_store, 0);
    }
  }
}

/* redodims: the original (or the default) first, then store buffers for the outputs this op creates. */
static pdl_error pdlb200_pp_redodims(Core *PDLc, pdl_trans *tr, pdlb200_trans_fn orig) {
  pdl_error PDL_err = orig ? orig(tr) : PDLc->redodims_default(tr);
  if (!PDL_err.error) pdlb200_pp_give_outputs(PDLc, tr);
  return PDL_err;
}

/* readdata: `fallback` (may be NULL) is the reference's own readdata for what is not on the device path. */
static pdl_error pdlb200_pp_readdata(Core *PDLc, pdl_trans *tr, int opid, pdlb200_trans_fn fallback) {
  pdl_error PDL_err = {0, NULL, 0};
  pdl_transvtable *vt = tr->vtable;
  pdlb200_trans d;
  pdlb200_pp_staged_t st[2 * PDLB200_MAXPDLS];
  int nst = 0, rc, need_sync = 0, on_device = pdlb200_pp_enabled;
  int32_t anybad = 0;
  PDL_Indx i, j, npdls = vt->npdls;
  char err[512];
  if (!pdlb200_pp_type_ok(tr->__datatype, opid) || npdls > PDLB200_MAXPDLS || tr->broadcast.ndims > PDLB200_MAXDIMS) on_device = 0;
  for (j = 0; on_device && j < npdls; j++)
    if (!pdlb200_pp_type_ok(tr->pdls[j]->datatype, opid) || !pdlb200_pp_cd_badval_ok(PDLc, tr->pdls[j])) on_device = 0;
  if (!on_device) {
    if (fallback) {
      /* the reference's own host loop: its parameters must be current in host memory */
      for (j = 0; j < npdls; j++) {
        pdl *p = tr->pdls[j], *owner = PDL_VAFFOK(p) ? p->vafftrans->from : p;
        if (owner->data && pdlb200_mbuf_host(owner->data, j >= vt->nparents, err, sizeof err) != 0)
          return PDLc->make_error(PDL_EFATAL, "PDL::B200 %s: %s", vt->name, err);
      }
      pdlb200_pp_host_calls++;
      return fallback(tr);
    }
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
  } else if (opid == PDLB200_OP_MINIMUM_N_IND || opid == PDLB200_OP_MAXIMUM_N_IND) {
    /* a(n); indx [o]c(m): ind_names sorted m, n; the ABI wants n, m; inc_sizes = {inc_a_n, inc_c_m} */
    d.ind[0] = tr->ind_sizes[1]; d.ind[1] = tr->ind_sizes[0];
    d.rinc[0] = tr->inc_sizes[0]; d.rinc[1] = tr->inc_sizes[1];
  } else if (vt->ninds >= 1) {
    d.ind[0] = tr->ind_sizes[0];
    for (i = 0; i < vt->nind_ids && i < 8; i++) d.rinc[i] = tr->inc_sizes[i];
  }
  /* OtherPars: setvaltobad(double value) / setbadtoval(double newval) — the params struct holds that one double
   * ([gen] Bad-pp-setvaltobad.c: typedef struct pdl_params_setvaltobad { double value; }) */
  if ((opid == PDLB200_OP_SETVALTOBAD || opid == PDLB200_OP_SETBADTOVAL) && tr->params) d.param = *(double *)tr->params;
  d.anybad = &anybad;
  for (j = 0; j < npdls; j++) {
    pdl *p = tr->pdls[j];
    int is_out = j >= vt->nparents;
    void *base = NULL;
    if (p->nvals > 0) {
      /* an output that is a window into a bigger buffer (vaffine) is only partly overwritten: keep the rest */
      base = pdlb200_pp_device_view(PDLc, p, !is_out || PDL_VAFFOK(p), is_out, st, &nst, &need_sync, err, sizeof err);
      if (!base) {
        pdlb200_pp_finish_staged(st, nst, 0, err + 400, 100);
        return PDLc->make_error(PDL_EFATAL, "PDL::B200 %s: no device-usable storage for parameter %s: %s", vt->name, vt->par_names[j], err);
      }
    }
    d.pdls[j].data = base;
    d.pdls[j].offs = PDL_REPROFFS(p);
    d.pdls[j].type = p->datatype;
    d.pdls[j].badval = pdlb200_pp_badval_bits(PDLc, p);
    d.pdls[j].flags = ((p->state & PDL_BADVAL) ? PDLB200_PAR_BADFLAG : 0) | (pdlb200_pp_badval_isnan(PDLc, p) ? PDLB200_PAR_BADNAN : 0);
  }
  rc = pdlb200_readdata(&d, err, sizeof err);
  /* no CPU fallback for the device type matrix: an unsupported shape (e.g. > 8 unmergeable broadcast dims) is an error */
  if (rc == 0) rc = pdlb200_pp_finish_staged(st, nst, need_sync, err, sizeof err);
  else pdlb200_pp_finish_staged(st, nst, 0, err + 400, 100);
  if (rc != 0) return PDLc->make_error(PDL_EUSERERROR, "PDL::B200 %s: %s", vt->name, err);
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
  /* minimum_n_ind / maximum_n_ind: $PDLSTATESETGOOD(c), then SETBAD if a slot could not be filled (Ufunc.pd:521-533) */
  if (opid == PDLB200_OP_MINIMUM_N_IND || opid == PDLB200_OP_MAXIMUM_N_IND) {
    if (anybad) tr->pdls[1]->state |= PDL_BADVAL; else tr->pdls[1]->state &= ~PDL_BADVAL;
  }
  pdlb200_pp_device_calls++;
  if (pdlb200_pp_verbose) fprintf(stderr, "PDL::B200 %s -> %s\n", vt->name, pdlb200_last_kernel());
  return PDL_err;
}

/* ---- flat parent -> child transformations of Core / Slices: converttypei (lib/PDL/Core/pdlconv.c:130-201) and
 * _clump_int (lib/PDL/Slices.pd:1373-1402).  Both read PARENT[i] and write CHILD[i] for i < nvals over PHYSICAL
 * ndarrays (writebackdata: the other way round), so the descriptor is one flat dim.  `reverse` = writebackdata. */
static pdl_error pdlb200_pp_flat(Core *PDLc, pdl_trans *tr, int opid, int reverse, pdlb200_trans_fn fallback) {
  pdl_error PDL_err = {0, NULL, 0};
  pdl *from = tr->pdls[reverse ? 1 : 0], *to = tr->pdls[reverse ? 0 : 1];
  pdlb200_trans d;
  pdlb200_pp_staged_t st[4];
  int nst = 0, need_sync = 0, rc, k;
  char err[512];
  pdl *pp[2];
  if (!pdlb200_pp_enabled || from->datatype > PDL_D || to->datatype > PDL_D || PDL_VAFFOK(from) || PDL_VAFFOK(to) ||
      pdlb200_pp_init() != 0) {
    for (k = 0; k < 2; k++) {
      pdl *p = tr->pdls[k], *owner = PDL_VAFFOK(p) ? p->vafftrans->from : p;
      if (owner->data && pdlb200_mbuf_host(owner->data, p == to, err, sizeof err) != 0)
        return PDLc->make_error(PDL_EFATAL, "PDL::B200 %s: %s", tr->vtable->name, err);
    }
    pdlb200_pp_host_calls++;
    return fallback(tr);
  }
  if (to->nvals <= 0) return PDL_err;
  if (from->data == to->data) return PDL_err;   /* child aliases the parent's buffer (clump of a physical parent): nothing to move */
  memset(&d, 0, sizeof d);
  d.op = (from->datatype == to->datatype) ? PDLB200_OP_ASSGN : PDLB200_OP_CONVERT;
  (void)opid;
  d.datatype = from->datatype;
  d.bvalflag = reverse ? ((from->state & PDL_BADVAL) ? 1 : 0) : (tr->bvalflag ? 1 : 0);
  d.npdls = 2; d.ndims = 1; d.dims[0] = to->nvals; d.incs[0] = 1; d.incs[1] = 1;
  pp[0] = from; pp[1] = to;
  for (k = 0; k < 2; k++) {
    void *base = pdlb200_pp_device_view(PDLc, pp[k], k == 0, k == 1, st, &nst, &need_sync, err, sizeof err);
    if (!base) { pdlb200_pp_finish_staged(st, nst, 0, err + 400, 100); return PDLc->make_error(PDL_EFATAL, "PDL::B200 %s: %s", tr->vtable->name, err); }
    d.pdls[k].data = base; d.pdls[k].offs = 0; d.pdls[k].type = pp[k]->datatype;
    d.pdls[k].badval = pdlb200_pp_badval_bits(PDLc, pp[k]);
    d.pdls[k].flags = ((pp[k]->state & PDL_BADVAL) ? PDLB200_PAR_BADFLAG : 0) | (pdlb200_pp_badval_isnan(PDLc, pp[k]) ? PDLB200_PAR_BADNAN : 0);
  }
  /* converttype maps BAD to the TARGET type's badvalue (pdlconv.c:87-95); the child of a BAD parent is flagged by core */
  if (d.bvalflag) d.pdls[1].flags |= PDLB200_PAR_BADFLAG;
  rc = pdlb200_readdata(&d, err, sizeof err);
  if (rc == 0) rc = pdlb200_pp_finish_staged(st, nst, need_sync, err, sizeof err);
  else pdlb200_pp_finish_staged(st, nst, 0, err + 400, 100);
  if (rc != 0) return PDLc->make_error(PDL_EUSERERROR, "PDL::B200 %s: %s", tr->vtable->name, err);
  pdlb200_pp_device_calls++;
  if (pdlb200_pp_verbose) fprintf(stderr, "PDL::B200 %s%s -> %s\n", tr->vtable->name, reverse ? " (writeback)" : "", pdlb200_last_kernel());
  return PDL_err;
}

/* redodims of _clump_int for a physical parent that already lives in the store: the child is the SAME bytes in
 * the same order, so it shares the parent's buffer (one more owner) instead of receiving a copy — the two-way
 * dataflow the reference implements with readdata + writebackdata copies becomes the identity. */
static void pdlb200_pp_alias_child(Core *PDLc, pdl_trans *tr) {
  pdl *par = tr->pdls[0], *ch = tr->pdls[1];
  if (!pdlb200_pp_enabled || !par || !ch || ch->data || (ch->state & PDL_ALLOCATED) || ch->trans_parent != tr) return;
  if (PDL_VAFFOK(par) || !par->data || !(par->state & PDL_ALLOCATED) || !pdlb200_mbuf_is(par->data)) return;
  if (par->datatype != ch->datatype || par->nvals != ch->nvals || par->has_badvalue || ch->has_badvalue) return;
  pdlb200_mbuf_retain(par->data);
  ch->data = par->data; ch->nbytes = par->nbytes;
  ch->state |= PDL_ALLOCATED | PDL_DONTTOUCHDATA;
  PDLc->add_deletedata_magic(ch, pdlb200_pp_free_store, 0);
}

/* ---- the Core function table: every transformation passes through PDL->make_trans_mutual (pdlapi.c:746-823)
 * before its readdata can run.  For the ones that are NOT device ops and are not pure index arithmetic (affine),
 * the parameters' data is made current in host memory first — the explicit form of the host-access choke point
 * "make_physical loop of every CPU op" (pdlapi.c:102-110); it also keeps pthreaded CPU loops from faulting. */
static pdl_error (*pdlb200_pp_orig_mtm)(pdl_trans *) = NULL;
static Core *pdlb200_pp_core = NULL;

static void pdlb200_pp_host_current(pdl *p, int for_write, int depth) {
  char err[256];
  if (!p || depth > 32) return;
  if (p->data && (p->state & PDL_ALLOCATED)) pdlb200_mbuf_host(p->data, for_write, err, sizeof err);
  if (PDL_VAFFOK(p)) { pdlb200_pp_host_current(p->vafftrans->from, for_write, depth + 1); return; }
  if (p->trans_parent && (p->trans_parent->flags & PDL_ITRANS_ISAFFINE) && p->trans_parent->pdls[0] != p)
    pdlb200_pp_host_current(p->trans_parent->pdls[0], for_write, depth + 1);
}

static pdl_error pdlb200_pp_make_trans_mutual(pdl_trans *tr) {
  pdl_transvtable *vt = tr->vtable;
  if (vt && !(vt->iflags & PDL_ITRANS_ISAFFINE) && !pdlb200_devop_is(vt)) {
    PDL_Indx j;
    for (j = 0; j < vt->npdls; j++) pdlb200_pp_host_current(tr->pdls[j], j >= vt->nparents, 0);
    pdlb200_pp_cpu_trans++;
  }
  return pdlb200_pp_orig_mtm(tr);
}

static void pdlb200_pp_hook_core(Core *PDLc) {
  if (pdlb200_pp_core) return;
  pdlb200_pp_core = PDLc;
  pdlb200_pp_orig_mtm = PDLc->make_trans_mutual;
  PDLc->make_trans_mutual = pdlb200_pp_make_trans_mutual;
}
static void pdlb200_pp_unhook_core(void) {
  if (pdlb200_pp_core && pdlb200_pp_core->make_trans_mutual == pdlb200_pp_make_trans_mutual)
    pdlb200_pp_core->make_trans_mutual = pdlb200_pp_orig_mtm;
  pdlb200_pp_core = NULL;
}

#endif /* PDLB200_PP_H */
