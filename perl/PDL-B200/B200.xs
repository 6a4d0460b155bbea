/*
 * PDL::B200 — attaches libpdlb200 (include/pdlb200.h) to an UNMODIFIED PDL at run time.
 *
 * The reference exports, for every pp_def, a `pdl_transvtable pdl_<op>_vtable` whose
 * `readdata` member core calls from pdl__ensure_trans (lib/PDL/Core/pdlapi.c:9-39,91-119).
 * attach() stores b200_readdata there (and b200_redodims in `redodims`), ORs in
 * PDL_TRANS_NO_PARALLEL so autopthread never re-enters us from worker threads
 * (lib/PDL/Core/pdlapi.c:842, pdlbroadcast.c:192), and keeps the original pointers: they
 * still serve the types that have no device representation (long double, complex).
 *
 * Data store (north-star subsystem 1) without touching Core: ndarray data lives in CUDA
 * managed memory.  b200_redodims gives every output the op creates a managed buffer BEFORE
 * core's PDL_ENSURE_ALLOCATED would allocate + zero-fill a Perl SV (pdlapi.c:14-17,172-209);
 * large host-backed inputs are moved into managed memory the first time a device op reads them
 * (their Perl SV is released, the pdl becomes PDL_DONTTOUCHDATA like an mmapped ndarray);
 * small ones (<= STAGE_MAX bytes, e.g. Perl scalars and 0-dim outputs stored inline in
 * pdl.value) go through a pinned staging buffer.  Host code keeps dereferencing pdl->data;
 * the driver migrates pages on demand (lazy host sync) and chained ops never cross PCIe.
 */
#define PERL_NO_GET_CONTEXT
#include "EXTERN.h"
#include "perl.h"
#include "XSUB.h"

#include "pdl.h"
#include "pdlcore.h"

#include "pdlb200.h"

static Core *PDL;

#define MAX_HOOKS 64
#define STAGE_MAX 65536              /* parameters up to this many bytes are staged, not migrated */
#define STAGE_BYTES (8 * STAGE_MAX)

typedef pdl_error (*trans_fn)(pdl_trans *);
typedef struct {
  pdl_transvtable *vt;
  trans_fn orig_readdata, orig_redodims;
  int opid, saved_flags;
} hook_t;

static hook_t hooks[MAX_HOOKS];
static int nhooks = 0;
static int g_enabled = 1;
static int g_verbose = 0;
static unsigned long g_device_calls = 0, g_host_calls = 0, g_migrated = 0, g_staged = 0;
static char *g_stage = NULL;         /* pinned, device-visible */
static size_t g_stage_used = 0;

static hook_t *find_hook(pdl_transvtable *vt) {
  int i;
  for (i = 0; i < nhooks; i++) if (hooks[i].vt == vt) return &hooks[i];
  return NULL;
}

static void b200_free_managed(pdl *it, Size_t param) {
  (void)param;
  if (it->data) { pdlb200_managed_free(it->data); it->data = NULL; }
}

static uint64_t badval_bits(pdl *p) {
  uint64_t bits = 0;
#define X(sym, ctype, ppsym, ...) \
  case sym: { ctype v = p->has_badvalue ? p->badvalue.value.ppsym : PDL->bvals.ppsym; memcpy(&bits, &v, sizeof v <= 8 ? sizeof v : 8); } break;
  switch (p->datatype) {
    PDL_TYPELIST_REAL(X)
    default: break;
  }
#undef X
  return bits;
}
static int badval_isnan(pdl *p) {
  if (p->datatype == PDL_F) { float v = p->has_badvalue ? p->badvalue.value.F : PDL->bvals.F; return v != v; }
  if (p->datatype == PDL_D) { double v = p->has_badvalue ? p->badvalue.value.D : PDL->bvals.D; return v != v; }
  return 0;
}

/* Move a physical ndarray's data into managed memory (once). */
static int migrate(pdl *it) {
  dTHX;
  void *m;
  if (!it->data || it->nbytes <= 0) return 0;
  m = pdlb200_managed_alloc((size_t)it->nbytes);
  if (!m) return -1;
  memcpy(m, it->data, (size_t)it->nbytes);
  if (it->datasv) { SvREFCNT_dec((SV *)it->datasv); it->datasv = NULL; }
  it->data = m;
  it->state |= PDL_DONTTOUCHDATA | PDL_ALLOCATED;
  PDL->add_deletedata_magic(it, b200_free_managed, 0);
  g_migrated++;
  return 0;
}

typedef struct { pdl *owner; char *slot; size_t nbytes; } staged_t;

/* Device-usable pointer for the buffer that physically holds p's data. */
static void *device_view(pdl *p, int is_output, staged_t *st, int *nst) {
  pdl *owner = PDL_VAFFOK(p) ? p->vafftrans->from : p;
  int kind;
  if (!owner->data) return NULL;
  kind = pdlb200_ptr_kind(owner->data);
  if (kind != 0) return owner->data;          /* managed, pinned or device memory already */
  if ((size_t)owner->nbytes <= STAGE_MAX && g_stage_used + (size_t)owner->nbytes + 64 <= STAGE_BYTES) {
    int i;
    char *slot;
    for (i = 0; i < *nst; i++)
      if (st[i].owner == owner) {               /* aliasing parameters (inplace ops) share a slot */
        if (is_output) st[i].nbytes = (size_t)owner->nbytes;
        return st[i].slot;
      }
    slot = g_stage + g_stage_used;
    g_stage_used += ((size_t)owner->nbytes + 63) & ~(size_t)63;
    memcpy(slot, owner->data, (size_t)owner->nbytes);
    st[*nst].owner = owner; st[*nst].slot = slot; st[*nst].nbytes = is_output ? (size_t)owner->nbytes : 0;
    (*nst)++;
    g_staged++;
    return slot;
  }
  if (migrate(owner) != 0) return NULL;
  return owner->data;
}

static pdl_error b200_redodims(pdl_trans *tr) {
  pdl_error PDL_err = {0, NULL, 0};
  hook_t *h = find_hook(tr->vtable);
  PDL_Indx i;
  PDL_err = (h && h->orig_redodims) ? h->orig_redodims(tr) : PDL->redodims_default(tr);
  if (PDL_err.error || !g_enabled || !h || tr->__datatype > PDL_D) return PDL_err;
  for (i = tr->vtable->nparents; i < tr->vtable->npdls; i++) {
    pdl *o = tr->pdls[i];
    PDL_Indx nbytes;
    if (!o || (o->state & PDL_ALLOCATED) || o->data || o->datatype > PDL_D || o->nvals <= 0) continue;
    if (o->trans_parent != tr) continue;       /* only ndarrays this op creates */
    nbytes = o->nvals * (PDL_Indx)PDL->howbig(o->datatype);
    if (nbytes <= STAGE_MAX) continue;         /* small outputs keep core's inline / SV storage */
    {
      void *m = pdlb200_managed_alloc((size_t)nbytes);
      if (!m) continue;                        /* core will allocate host memory; we migrate later */
      o->data = m; o->nbytes = nbytes;
      o->state |= PDL_ALLOCATED | PDL_DONTTOUCHDATA;
      PDL->add_deletedata_magic(o, b200_free_managed, 0);
    }
  }
  return PDL_err;
}

static pdl_error b200_readdata(pdl_trans *tr) {
  pdl_error PDL_err = {0, NULL, 0};
  hook_t *h = find_hook(tr->vtable);
  pdl_transvtable *vt = tr->vtable;
  pdlb200_trans d;
  staged_t st[PDLB200_MAXPDLS];
  int nst = 0, rc;
  PDL_Indx i, j, npdls = vt->npdls;
  char err[512];
  if (!h) return PDL->make_error_simple(PDL_EFATAL, "PDL::B200: readdata called for an unhooked vtable");
  if (!g_enabled || tr->__datatype > PDL_D || npdls > PDLB200_MAXPDLS || tr->broadcast.ndims > PDLB200_MAXDIMS) {
    g_host_calls++;
    return h->orig_readdata(tr);
  }
  for (j = 0; j < npdls; j++)
    if (tr->pdls[j]->datatype > PDL_D) { g_host_calls++; return h->orig_readdata(tr); }

  memset(&d, 0, sizeof d);
  d.op = h->opid; d.datatype = tr->__datatype; d.bvalflag = tr->bvalflag;
  d.npdls = (int32_t)npdls; d.ndims = (int32_t)tr->broadcast.ndims;
  for (i = 0; i < tr->broadcast.ndims; i++) {
    d.dims[i] = tr->broadcast.dims[i];
    for (j = 0; j < npdls; j++) d.incs[i * npdls + j] = PDL_BRC_INC(tr->broadcast.incs, npdls, j, i);
  }
  if (h->opid == PDLB200_OP_MATMULT) {
    /* ind_sizes are sorted by name: h, t, w ([gen] Primitive-pp-matmult.c); the ABI wants t, h, w */
    d.ind[0] = tr->ind_sizes[1]; d.ind[1] = tr->ind_sizes[0]; d.ind[2] = tr->ind_sizes[2];
    for (i = 0; i < 6; i++) d.rinc[i] = tr->inc_sizes[i];
  } else if (vt->ninds >= 1) {
    d.ind[0] = tr->ind_sizes[0];
    for (i = 0; i < vt->nind_ids && i < 8; i++) d.rinc[i] = tr->inc_sizes[i];
  }
  g_stage_used = 0;
  for (j = 0; j < npdls; j++) {
    pdl *p = tr->pdls[j];
    int is_out = j >= vt->nparents;
    void *base = NULL;
    if (p->nvals > 0) {
      base = device_view(p, is_out, st, &nst);
      if (!base) return PDL->make_error(PDL_EFATAL, "PDL::B200 %s: no device-usable storage for parameter %s", vt->name, vt->par_names[j]);
    }
    d.pdls[j].data = base;
    d.pdls[j].offs = PDL_REPROFFS(p);
    d.pdls[j].type = p->datatype;
    d.pdls[j].badval = badval_bits(p);
    d.pdls[j].flags = ((p->state & PDL_BADVAL) ? PDLB200_PAR_BADFLAG : 0) | (badval_isnan(p) ? PDLB200_PAR_BADNAN : 0);
  }
  rc = pdlb200_readdata(&d, err, sizeof err);
  if (rc == PDLB200_EUNSUPPORTED) { g_host_calls++; return h->orig_readdata(tr); }  /* e.g. > 8 unmergeable dims */
  if (rc == 0) rc = pdlb200_sync(NULL, err, sizeof err);   /* host code may read pdl->data as soon as we return */
  if (rc != 0) return PDL->make_error(PDL_EUSERERROR, "PDL::B200 %s: %s", vt->name, err);
  for (i = 0; i < nst; i++)
    if (st[i].nbytes) memcpy(st[i].owner->data, st[i].slot, st[i].nbytes);
  /* outputs flagged BAD by the op itself: minimum/maximum(_ind) with no good element (Ufunc.pd:463-464) */
  if (h->opid >= PDLB200_OP_MINIMUM && h->opid <= PDLB200_OP_MAXIMUM_IND && !tr->bvalflag && tr->ind_sizes[0] == 0)
    tr->pdls[1]->state |= PDL_BADVAL;
  g_device_calls++;
  if (g_verbose) fprintf(stderr, "PDL::B200 %s -> %s\n", vt->name, pdlb200_last_kernel());
  return PDL_err;
}

MODULE = PDL::B200   PACKAGE = PDL::B200

PROTOTYPES: DISABLE

BOOT:
{
  SV *CoreSV;
  perl_require_pv("PDL/Core.pm");
  if (SvTRUE(ERRSV)) Perl_croak(aTHX_ "%s", SvPV_nolen(ERRSV));
  CoreSV = perl_get_sv("PDL::SHARE", FALSE);
  if (!CoreSV) Perl_croak(aTHX_ "PDL::B200 requires the PDL::Core module, which was not found");
  if (!(PDL = INT2PTR(Core *, SvIV(CoreSV)))) Perl_croak(aTHX_ "Got NULL pointer for PDL");
  if (PDL->Version != PDL_CORE_VERSION)
    Perl_croak(aTHX_ "[PDL->Version: %ld PDL_CORE_VERSION: %ld] PDL::B200 needs to be recompiled against the installed PDL",
               (long)PDL->Version, (long)PDL_CORE_VERSION);
}

int
device_count()
CODE:
  RETVAL = pdlb200_device_count();
OUTPUT:
  RETVAL

int
_hook(vtable_addr, opid)
  IV vtable_addr
  int opid
CODE:
  {
    pdl_transvtable *vt = INT2PTR(pdl_transvtable *, vtable_addr);
    hook_t *h = find_hook(vt);
    if (pdlb200_device_count() <= 0)
      Perl_croak(aTHX_ "PDL::B200: no CUDA device available and there is no CPU fallback to attach");
    if (!g_stage) {
      g_stage = (char *)pdlb200_host_alloc(STAGE_BYTES);
      if (!g_stage) Perl_croak(aTHX_ "PDL::B200: cannot allocate the pinned staging buffer");
    }
    if (!h) {
      if (nhooks >= MAX_HOOKS) Perl_croak(aTHX_ "PDL::B200: too many hooks");
      h = &hooks[nhooks++];
      h->vt = vt; h->opid = opid;
      h->orig_readdata = vt->readdata; h->orig_redodims = vt->redodims; h->saved_flags = vt->flags;
    }
    vt->readdata = b200_readdata;
    vt->redodims = b200_redodims;
    vt->flags |= PDL_TRANS_NO_PARALLEL;
    RETVAL = nhooks;
  }
OUTPUT:
  RETVAL

void
detach()
CODE:
  {
    int i;
    for (i = 0; i < nhooks; i++) {
      hooks[i].vt->readdata = hooks[i].orig_readdata;
      hooks[i].vt->redodims = hooks[i].orig_redodims;
      hooks[i].vt->flags = hooks[i].saved_flags;
    }
    nhooks = 0;
  }

void
enable(on)
  int on
CODE:
  g_enabled = on;

void
verbose(on)
  int on
CODE:
  g_verbose = on;

void
stats()
PPCODE:
  EXTEND(SP, 5);
  mPUSHu(g_device_calls);
  mPUSHu(g_host_calls);
  mPUSHu(g_migrated);
  mPUSHu(g_staged);
  mPUSHu((UV)pdlb200_launch_count());

const char *
last_kernel()
CODE:
  RETVAL = pdlb200_last_kernel();
OUTPUT:
  RETVAL

int
ptr_kind(p)
  pdl *p
CODE:
  {
    pdl *owner = PDL_VAFFOK(p) ? p->vafftrans->from : p;
    RETVAL = owner->data ? pdlb200_ptr_kind(owner->data) : -1;
  }
OUTPUT:
  RETVAL

void
to_device(p)
  pdl *p
CODE:
  {
    /* make the ndarray's data managed and prefetch it into HBM */
    char err[256];
    pdl *owner;
    pdl_error e = PDL->make_physvaffine(p);
    if (e.error) PDL->pdl_barf("PDL::B200::to_device: make_physvaffine failed");
    owner = PDL_VAFFOK(p) ? p->vafftrans->from : p;
    if (owner->data && pdlb200_ptr_kind(owner->data) == 0 && migrate(owner) != 0)
      PDL->pdl_barf("PDL::B200::to_device: managed allocation failed");
    if (owner->data && pdlb200_prefetch(owner->data, (size_t)owner->nbytes, 1, NULL, err, sizeof err) != 0)
      PDL->pdl_barf("PDL::B200::to_device: %s", err);
    pdlb200_sync(NULL, err, sizeof err);
  }
