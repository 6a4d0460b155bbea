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

#include "pdlb200_pp.h"

static Core *PDL;

#define MAX_HOOKS 128
typedef struct {
  pdl_transvtable *vt;
  pdlb200_trans_fn orig_readdata, orig_redodims;
  int opid, saved_flags;
} hook_t;

static hook_t hooks[MAX_HOOKS];
static int nhooks = 0;

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
    if (pdlb200_pp_init() != 0) Perl_croak(aTHX_ "PDL::B200: cannot allocate the pinned staging buffer");
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
  pdlb200_pp_enabled = on;

void
verbose(on)
  int on
CODE:
  pdlb200_pp_verbose = on;

void
stats()
PPCODE:
  EXTEND(SP, 5);
  mPUSHu(pdlb200_pp_device_calls);
  mPUSHu(pdlb200_pp_host_calls);
  mPUSHu(pdlb200_pp_migrated);
  mPUSHu(pdlb200_pp_staged);
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
    if (owner->data && pdlb200_ptr_kind(owner->data) == 0 && pdlb200_pp_migrate(PDL, owner) != 0)
      PDL->pdl_barf("PDL::B200::to_device: managed allocation failed");
    if (owner->data && pdlb200_prefetch(owner->data, (size_t)owner->nbytes, 1, NULL, err, sizeof err) != 0)
      PDL->pdl_barf("PDL::B200::to_device: %s", err);
    pdlb200_sync(NULL, err, sizeof err);
  }
