/*
 * PDL::B200 — attaches libpdlb200 (include/pdlb200.h) to an UNMODIFIED PDL at run time.
 *
 * The reference exports, for every pp_def, a `pdl_transvtable pdl_<op>_vtable` whose
 * `readdata` member core calls from pdl__ensure_trans (lib/PDL/Core/pdlapi.c:9-39,91-119).
 * attach() stores b200_readdata there (and b200_redodims in `redodims`), ORs in
 * PDL_TRANS_NO_PARALLEL so autopthread never re-enters us from worker threads
 * (lib/PDL/Core/pdlapi.c:842, pdlbroadcast.c:192), and keeps the original pointers: they
 * still serve the types that have no device representation (long double, complex).
 * Two transformations of Core / Slices that sit between device ops in ordinary scripts are hooked the same
 * way: converttypei (`float_nd + 1.5`, lib/PDL/Core/pdlconv.c:130-201, inserted by pdlapi.c:1262-1311) and
 * _clump_int (`$x->sum` = flat->sumover, lib/PDL/Slices.pd:1373-1402).
 *
 * Data store (north-star subsystem 1) without touching Core: see pdlb200_pp.h — cudaMalloc'd device buffers
 * with protected host mirrors and dirty bits (the library's pdlb200_mbuf_* store), no synchronisation between
 * chained device ops, host access made current at the Core function table and by the store's fault handler.
 */
#define PERL_NO_GET_CONTEXT
#include "EXTERN.h"
#include "perl.h"
#include "XSUB.h"

#include "pdl.h"
#include "pdlcore.h"

#include "pdlb200_pp.h"

static Core *PDL;

#define MAX_HOOKS 128
typedef struct {
  pdl_transvtable *vt;
  pdlb200_trans_fn orig_readdata, orig_redodims, orig_writeback;
  int opid, saved_flags, flat;   /* flat: 1 converttypei, 2 _clump_int */
} hook_t;

static hook_t hooks[MAX_HOOKS];
static int nhooks = 0;
static pdl_transvtable *mult_vtable = NULL;

static hook_t *find_hook(pdl_transvtable *vt) {
  int i;
  for (i = 0; i < nhooks; i++) if (hooks[i].vt == vt) return &hooks[i];
  return NULL;
}

static pdl_error b200_redodims(pdl_trans *tr) {
  hook_t *h = find_hook(tr->vtable);
  return pdlb200_pp_redodims(PDL, tr, h ? h->orig_redodims : NULL);
}

static pdl_error b200_readdata(pdl_trans *tr) {
  hook_t *h = find_hook(tr->vtable);
  if (!h) return PDL->make_error_simple(PDL_EFATAL, "PDL::B200: readdata called for an unhooked vtable");
  return pdlb200_pp_readdata(PDL, tr, h->opid, h->orig_readdata);
}

/* converttypei / _clump_int */
static pdl_error b200_flat_redodims(pdl_trans *tr) {
  hook_t *h = find_hook(tr->vtable);
  pdl_error e = h->orig_redodims(tr);
  if (e.error) return e;
  if (h->flat == 2) pdlb200_pp_alias_child(PDL, tr);
  pdlb200_pp_give_outputs(PDL, tr);
  return e;
}
static pdl_error b200_flat_readdata(pdl_trans *tr) {
  hook_t *h = find_hook(tr->vtable);
  return pdlb200_pp_flat(PDL, tr, h->opid, 0, h->orig_readdata);
}
static pdl_error b200_flat_writeback(pdl_trans *tr) {
  hook_t *h = find_hook(tr->vtable);
  return pdlb200_pp_flat(PDL, tr, h->opid, 1, h->orig_writeback);
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
_hook(vtable_addr, opid, flat = 0)
  IV vtable_addr
  int opid
  int flat
CODE:
  {
    pdl_transvtable *vt = INT2PTR(pdl_transvtable *, vtable_addr);
    hook_t *h = find_hook(vt);
    if (pdlb200_device_count() <= 0)
      Perl_croak(aTHX_ "PDL::B200: no CUDA device available and there is no CPU fallback to attach");
    if (pdlb200_pp_init() != 0) Perl_croak(aTHX_ "PDL::B200: cannot allocate the pinned staging buffer");
    if (!h) {
      if (nhooks >= MAX_HOOKS) Perl_croak(aTHX_ "PDL::B200: too many hooks");
      h = &hooks[nhooks++];
      h->vt = vt; h->opid = opid; h->flat = flat;
      h->orig_readdata = vt->readdata; h->orig_redodims = vt->redodims; h->orig_writeback = vt->writebackdata;
      h->saved_flags = vt->flags;
    }
    if (flat) {
      vt->readdata = b200_flat_readdata;
      vt->redodims = b200_flat_redodims;
      if (vt->writebackdata) vt->writebackdata = b200_flat_writeback;
    } else {
      vt->readdata = b200_readdata;
      vt->redodims = b200_redodims;
      vt->flags |= PDL_TRANS_NO_PARALLEL;
    }
    if (opid == PDLB200_OP_MULT) mult_vtable = vt;
    pdlb200_devop_register(vt, 1);
    pdlb200_pp_hook_core(PDL);
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
      if (hooks[i].flat) hooks[i].vt->writebackdata = hooks[i].orig_writeback;
      hooks[i].vt->flags = hooks[i].saved_flags;
      pdlb200_devop_register(hooks[i].vt, 0);
    }
    nhooks = 0;
    mult_vtable = NULL;
    pdlb200_pp_unhook_core();
  }

void
enable(on)
  int on
CODE:
  pdlb200_pp_enabled = on;

void
verbose(on)
  int on
CODE:
  pdlb200_pp_verbose = on;

void
stats()
PPCODE:
  /* device readdata calls, host (fallback) calls, ndarrays adopted into the store, staged parameters, kernels
   * launched, stream synchronisations issued by the binding, parameters staged through a temporary device buffer,
   * CPU transformations seen in the Core function table */
  EXTEND(SP, 8);
  mPUSHu(pdlb200_pp_device_calls);
  mPUSHu(pdlb200_pp_host_calls);
  mPUSHu(pdlb200_pp_migrated);
  mPUSHu(pdlb200_pp_staged);
  mPUSHu((UV)pdlb200_launch_count());
  mPUSHu(pdlb200_pp_syncs);
  mPUSHu(pdlb200_pp_transient);
  mPUSHu(pdlb200_pp_cpu_trans);

void
store_stats()
PPCODE:
  {
    /* buffers created, recycled, uploads, upload bytes, downloads, download bytes, faults handled, adopted */
    uint64_t s[8];
    int i;
    pdlb200_mbuf_stats(s);
    EXTEND(SP, 8);
    for (i = 0; i < 8; i++) mPUSHu((UV)s[i]);
  }

const char *
last_kernel()
CODE:
  RETVAL = pdlb200_last_kernel();
OUTPUT:
  RETVAL

int
store_state(p)
  pdl *p
CODE:
  {
    /* -1: plain host memory; else bits 0-1 host mirror (0 stale, 1 current, 2 current + host-modified), bit 2 device copy current */
    pdl *owner = PDL_VAFFOK(p) ? p->vafftrans->from : p;
    RETVAL = owner->data ? pdlb200_mbuf_state(owner->data) : -1;
  }
OUTPUT:
  RETVAL

void
to_device(p)
  pdl *p
CODE:
  {
    /* adopt the ndarray's data into the store now (one H2D copy) instead of at its first device op */
    char err[256];
    pdl *owner;
    pdl_error e = PDL->make_physvaffine(p);
    if (e.error) PDL->pdl_barf("PDL::B200::to_device: make_physvaffine failed");
    owner = PDL_VAFFOK(p) ? p->vafftrans->from : p;
    if (pdlb200_pp_init() != 0) PDL->pdl_barf("PDL::B200::to_device: no CUDA device");
    if (owner->data && !pdlb200_mbuf_is(owner->data) && pdlb200_pp_adoptable(owner) &&
        pdlb200_pp_adopt(PDL, owner, 1, err, sizeof err) != 0)
      PDL->pdl_barf("PDL::B200::to_device: %s", err);
    if (owner->data && pdlb200_mbuf_is(owner->data) && !pdlb200_mbuf_dev(owner->data, 0, 0, NULL, err, sizeof err))
      PDL->pdl_barf("PDL::B200::to_device: %s", err);
  }

void
to_host(p)
  pdl *p
CODE:
  {
    /* hand the ndarray back to plain host memory (a Perl SV): for get_dataref / setdims / reshape, which
     * refuse PDL_DONTTOUCHDATA ndarrays */
    if (p->data && (p->state & PDL_ALLOCATED) && pdlb200_pp_release(PDL, p) != 0)
      PDL->pdl_barf("PDL::B200::to_host: download failed");
  }

void
sync()
CODE:
  {
    char err[256];
    if (pdlb200_sync(NULL, err, sizeof err) != 0) PDL->pdl_barf("PDL::B200::sync: %s", err);
  }

void
_pending_product(c)
  pdl *c
PPCODE:
  {
    /* `($a->flowing * $b)`: the product's readdata is still deferred (PDL_ITRANS_DO_DATAFLOW_F,
     * lib/PDL/Core/pdlapi.c:781-801) and nothing else consumes it: return its two parents, so that a following
     * sumover can run fused with it (inner, lib/PDL/Primitive.pd:48-70).  Only where the fused form has the
     * unfused one's semantics: float/double, no BAD values. */
    pdl_trans *tr = c->trans_parent;
    if (tr && mult_vtable && tr->vtable == mult_vtable && (c->state & PDL_PARENTDATACHANGED) &&
        (tr->flags & PDL_ITRANS_DO_DATAFLOW_F) && c->ntrans_children == 0 &&
        (c->datatype == PDL_F || c->datatype == PDL_D) &&
        !(tr->pdls[0]->state & PDL_BADVAL) && !(tr->pdls[1]->state & PDL_BADVAL) &&
        tr->pdls[0]->datatype == c->datatype && tr->pdls[1]->datatype == c->datatype) {
      int k;
      EXTEND(SP, 2);
      for (k = 0; k < 2; k++) {
        SV *sv = sv_newmortal();
        PDL->SetSV_PDL(sv, tr->pdls[k]);
        PUSHs(sv);
      }
    }
  }
