/*
 * pdlb200.h — C-ABI of libpdlb200: the B200 (sm_100a) implementation of PDL's
 * PP-generated broadcast-loop hot path (PDL::Ops elementwise ops, PDL::Ufunc
 * reductions, PDL::Primitive::matmult).
 *
 * The boundary this library replaces is the reference's
 *     pdl_error (*readdata)(pdl_trans *)            lib/PDL/Core/pdl.h.PL:381-403
 * called from pdl__ensure_trans                      lib/PDL/Core/pdlapi.c:9-39,91-119
 * One call of pdlb200_readdata() == one call of a generated pdl_<op>_readdata().
 * The caller (an XS/PP shim that includes pdl.h; or the
 * Python mirror in pdl_b200/) copies the fields of pdl_trans it has already
 * computed (type_coerce + redodims done, every parameter physvaffine) into the
 * POD descriptor below.  Nothing here includes Perl, PDL or torch headers.
 *
 * Memory model: every `data` pointer is a DEVICE pointer (cudaMalloc'd, or any
 * allocation the CUDA context can address).  The library never frees or
 * reallocates a parameter buffer.  Launches are asynchronous on `stream`.
 * There is no CPU fallback: without a usable GPU every compute entry point
 * returns PDLB200_ENODEVICE and fills the error buffer.
 */
#ifndef PDLB200_H
#define PDLB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PDLB200_API __attribute__((visibility("default")))
#else
#define PDLB200_API
#endif

#define PDLB200_ABI_VERSION 4

/* Element types: numeric values are PDL's own pdl_datatypes enum
 * (lib/PDL/Types.pm:27-255, order is significant for promotion).
 * LD/CLD (11, 14) are x87 80-bit and have no GPU representation. */
enum {
  PDLB200_SB = 0, PDLB200_B = 1, PDLB200_S = 2, PDLB200_US = 3, PDLB200_L = 4,
  PDLB200_UL = 5, PDLB200_IND = 6, PDLB200_ULL = 7, PDLB200_LL = 8,
  PDLB200_F = 9, PDLB200_D = 10,
  PDLB200_NTYPES = 11,
  /* Complex float / double (C99 `float complex` / `double complex`, interleaved re, im): on the device path for
   * plus minus mult divide only (lib/PDL/Ops.pd:104-153,288-291).  mult is the compiler's inline
   * (ac-bd, ad+bc) with libgcc's __mulsc3/__muldc3 recovery of infinities, divide is libgcc's __divsc3 (in double)
   * / __divdc3 (scaled Smith), restated operation for operation.  `badval`: CF carries both parts (re in the low
   * 4 bytes); CD carries the bits of the REAL part and requires the imaginary part of the badvalue to be equal
   * (true of the default -DBL_MAX - DBL_MAX*i). */
  PDLB200_LD = 11, PDLB200_CF = 12, PDLB200_CD = 13, PDLB200_CLD = 14
};

/* Operations.  Each names the reference pp_def whose readdata it replaces. */
enum {
  /* biop(), lib/PDL/Ops.pd:104-153,288-313 : a(); b(); [o]c() */
  PDLB200_OP_PLUS = 0, PDLB200_OP_MULT, PDLB200_OP_MINUS, PDLB200_OP_DIVIDE,
  PDLB200_OP_GT, PDLB200_OP_LT, PDLB200_OP_LE, PDLB200_OP_GE, PDLB200_OP_EQ, PDLB200_OP_NE,
  PDLB200_OP_SHIFTLEFT, PDLB200_OP_SHIFTRIGHT, PDLB200_OP_OR2, PDLB200_OP_AND2, PDLB200_OP_XOR,
  /* bifunc(), lib/PDL/Ops.pd:164-219,321-324 : a(); b(); [o]c() */
  PDLB200_OP_POWER = 15, PDLB200_OP_ATAN2, PDLB200_OP_MODULO, PDLB200_OP_SPACESHIP,
  /* ufunc() and friends, lib/PDL/Ops.pd:222-265,327-397,491-503 : a(); [o]b() */
  PDLB200_OP_BITNOT = 19, PDLB200_OP_SQRT, PDLB200_OP_SIN, PDLB200_OP_COS, PDLB200_OP_NOT,
  PDLB200_OP_EXP, PDLB200_OP_LOG, PDLB200_OP_LOG10, PDLB200_OP_RABS, PDLB200_OP_ASSGN,
  PDLB200_OP_ABS2,
  /* reductions, lib/PDL/Ufunc.pd:88-118,413-444,446-500 : a(n); [o]b() */
  PDLB200_OP_SUMOVER = 30, PDLB200_OP_PRODOVER, PDLB200_OP_DSUMOVER, PDLB200_OP_DPRODOVER,
  PDLB200_OP_AVERAGE, PDLB200_OP_DAVERAGE, PDLB200_OP_MINIMUM, PDLB200_OP_MAXIMUM,
  PDLB200_OP_MINIMUM_IND, PDLB200_OP_MAXIMUM_IND,
  /* more a(n); [o]b() reductions, lib/PDL/Ufunc.pd:143-187 */
  PDLB200_OP_ANDOVER = 40, PDLB200_OP_OROVER, PDLB200_OP_BANDOVER, PDLB200_OP_BOROVER,
  PDLB200_OP_ZCOVER, PDLB200_OP_XOROVER, PDLB200_OP_BXOROVER,
  /* lib/PDL/Bad.pd:418-480 : a(n); indx [o]b() — the good/bad counts a sharded average needs */
  PDLB200_OP_NBADOVER = 47, PDLB200_OP_NGOODOVER,
  /* scans, lib/PDL/Ufunc.pd:120-141 : a(n); [o]b(n) */
  PDLB200_OP_CUMUSUMOVER = 50, PDLB200_OP_CUMUPRODOVER, PDLB200_OP_DCUMUSUMOVER, PDLB200_OP_DCUMUPRODOVER,
  /* matmult, lib/PDL/Primitive.pd:191-264 : a(t,h); b(w,t); [o]c(w,h) */
  PDLB200_OP_MATMULT = 60,
  /* type conversion, lib/PDL/Core/pdlconv.c:45-126,163-201 (converttypei) : a(); [o]b() of another type */
  PDLB200_OP_CONVERT = 61,
  /* ipow, lib/PDL/Ops.pd:443-476 : a(); longlong b(); [o]ans() — exponentiation by squaring */
  PDLB200_OP_IPOW = 62,
  /* bad-value producers/consumers either side of the path, lib/PDL/Bad.pd:343-416,584-905.
   * isbad/isgood/isnan: a(); int [o]b().  setbadif: a(); int mask(); [o]b().  setvaltobad/setbadtoval:
   * a(); [o]b() with the OtherPars double in `param`.  set{nan,inf,nonfinite}tobad, setbadtonan: a(); [o]b(),
   * floating point only.  badmask: a(); b(); [o]c().  copybad: a(); mask(); [o]b(). */
  PDLB200_OP_ISBAD = 63, PDLB200_OP_ISGOOD, PDLB200_OP_ISNAN, PDLB200_OP_SETBADIF, PDLB200_OP_SETVALTOBAD,
  PDLB200_OP_SETNANTOBAD, PDLB200_OP_SETINFTOBAD, PDLB200_OP_SETNONFINITETOBAD, PDLB200_OP_SETBADTONAN,
  PDLB200_OP_SETBADTOVAL, PDLB200_OP_BADMASK, PDLB200_OP_COPYBAD,
  /* axisvalues, lib/PDL/Primitive.pd:1468-1474 : i(n); [o]a(n) — a = n; the body of sequence/xvals/yvals/zvals
   * (lib/PDL/Basic.pm:117-129,479-485).  n_size = ind[0], inc_a_n = rinc[1] (rinc[0] is i's, never read). */
  PDLB200_OP_AXISVALUES = 75,
  /* inner, lib/PDL/Primitive.pd:48-70 : a(n); b(n); [o]c() — c = sum_n a*b; n_size = ind[0],
   * inc_a_n = rinc[0], inc_b_n = rinc[1].  Any BAD element in the row makes c BAD. */
  PDLB200_OP_INNER = 76,
  /* minmaximum, lib/PDL/Ufunc.pd:563-613 : a(n); [o]cmin(); [o]cmax(); indx [o]cmin_ind(); indx [o]cmax_ind()
   * (the body of `minmax`, Ufunc.pd:738).  BAD and NaN elements are skipped; a row without any other element
   * writes BAD to all four outputs and marks them BAD: `anybad` (required) reports that. */
  PDLB200_OP_MINMAXIMUM = 77,
  /* magnover, lib/PDL/Ufunc.pd:1235-1256 : a(n); real [o]b() — sqrt(sum_n a*a), BAD elements skipped */
  PDLB200_OP_MAGNOVER = 78,
  /* outer, lib/PDL/Primitive.pd:78-96 : a(n); b(m); [o]c(n,m) — c = a*b; ind = {n, m},
   * rinc = {inc_a_n, inc_b_m, inc_c_n, inc_c_m}.  Runs as mult over two extra leading broadcast dims. */
  PDLB200_OP_OUTER = 79,
  /* Whole-array reductions of an ndarray PARTITIONED across GPUs along its outermost dim (SURVEY.md §8(e);
   * the reference's own split rule, lib/PDL/Core/pdlbroadcast.c:469-484) collapse the sharded dim in three
   * steps, all on the device: (1) PART_*: the rank's block reduced to one 32-byte record per row,
   * a(n); longlong [o]rec(r=4) with rec = {value bits, good count, GLOBAL index, state}; state 0 = no good
   * element, 1 = value is not NaN, 2 = every good element is NaN (value = the LAST one, Ufunc.pd:455-465).
   * ind[0] = n, ind[1] = global index of this block's element 0, rinc[0] = inc_a_n, rinc[1] = inc_rec_r.
   * PART_SUM accumulates in the `int+` type of `datatype` (sumover/average, Ufunc.pd:88-118,413-444),
   * PART_DSUM in double (dsumover/daverage).  (2) the records of all ranks are gathered (ncclAllGather or
   * peer stores).  (3) COLL_*: longlong rec(r=4,k); [o]b() merges the k records of a row IN RANK ORDER with
   * the reference's rules (BAD skipped, all BAD -> BAD, NaN loses to non-NaN, first index wins), so every
   * rank holds the same bits.  `datatype` = the type of the VALUE in the records (the int+/double type for
   * SUM/AVG, the input type for MIN/MAX); ind[0] = k, rinc[0] = inc_rec_r, rinc[1] = inc_rec_k. */
  PDLB200_OP_PART_SUM = 80, PDLB200_OP_PART_DSUM, PDLB200_OP_PART_MIN, PDLB200_OP_PART_MAX,
  PDLB200_OP_COLL_SUM = 84, PDLB200_OP_COLL_AVG, PDLB200_OP_COLL_MIN, PDLB200_OP_COLL_MAX,
  PDLB200_OP_COLL_MIN_IND, PDLB200_OP_COLL_MAX_IND,
  /* minimum_n_ind / maximum_n_ind, lib/PDL/Ufunc.pd:502-561 : a(n); indx [o]c(m) — indices of the first m extreme
   * elements (m selection passes with minimum_ind's rule).  ind = {n, m}, rinc = {inc_a_n, inc_c_m}.  Slots that
   * cannot be filled are BAD and flag the output: `anybad` (required) reports that; otherwise the output's
   * badflag is CLEARED ($PDLSTATESETGOOD, Ufunc.pd:521). */
  PDLB200_OP_MINIMUM_N_IND = 90, PDLB200_OP_MAXIMUM_N_IND = 91,
  PDLB200_OP__END
};

#define PDLB200_MAXDIMS 16  /* broadcast dims carried per call (pdl_broadcast.ndims) */
#define PDLB200_MAXPDLS 5   /* parameters per transformation on this path (minmaximum has 5) */

/* pdlb200_trans.tflags */
#define PDLB200_TRANS_DEFER_ANYBAD 1 /* `anybad` points to PINNED, device-mapped host memory (pdlb200_host_alloc): the
                                      * call does NOT synchronise the stream — the flag arrives by a 4-byte async copy
                                      * or by a store from the kernel itself; the int32 is valid once the stream has
                                      * been synchronised past this call */

/* pdlb200_par.flags */
#define PDLB200_PAR_BADFLAG 1  /* pdl->state & PDL_BADVAL                  pdl.h.PL:560-561 */
#define PDLB200_PAR_BADNAN  2  /* the parameter's badvalue is NaN          pdlcore.h:202-205 */

/* One parameter (ndarray) of the transformation, after make_physvaffine:
 * element (i0,i1,...; n) lives at  data + (offs + sum_k i_k*incs[k][p] + n*inc_n) * sizeof(type)
 * (lib/PDL/Core/pdlbroadcast.h:64, pdl.h.PL:511-517 PDL_REPRP/PDL_REPROFFS). */
typedef struct pdlb200_par {
  void    *data;    /* device pointer: PDL_REPRP(pdl) */
  int64_t  offs;    /* PDL_REPROFFS(pdl), in elements */
  uint64_t badval;  /* bit pattern of this parameter's badvalue, in `type`, low-order bytes */
  int32_t  type;    /* PDLB200_* element type of this parameter */
  int32_t  flags;   /* PDLB200_PAR_* */
} pdlb200_par;

/* POD image of the parts of pdl_trans + pdl_broadcast a readdata reads
 * (lib/PDL/Core/pdl.h.PL:381-403,471-482; lib/PDL/Core/pdlbroadcast.h:18-37). */
typedef struct pdlb200_trans {
  int32_t op;        /* PDLB200_OP_* */
  int32_t datatype;  /* trans->__datatype: the generic type the loop is instantiated for */
  int32_t bvalflag;  /* trans->bvalflag: any input had PDL_BADVAL (pdlapi.c:760-766) */
  int32_t npdls;     /* vtable->npdls */
  int32_t ndims;     /* broadcast.ndims (0 allowed) */
  int32_t tflags;    /* PDLB200_TRANS_* (0 = the plain synchronous call) */
  int64_t dims[PDLB200_MAXDIMS];                    /* broadcast.dims */
  int64_t incs[PDLB200_MAXDIMS * PDLB200_MAXPDLS];  /* broadcast.incs[d*npdls + p], elements */
  /* Named ("real") dims.  reductions/scans: n_size = ind[0], inc_a_n = rinc[0],
   * (scans: inc_b_n = rinc[1]).  matmult: ind = {t, h, w} sizes and
   * rinc = {inc_a_t, inc_a_h, inc_b_w, inc_b_t, inc_c_w, inc_c_h}
   * (the generated code reads them as ind_sizes[]/inc_sizes[PDL_INC_ID()]). */
  int64_t ind[4];
  int64_t rinc[8];
  pdlb200_par pdls[PDLB200_MAXPDLS];                /* inputs first, then outputs */
  void   *stream;    /* cudaStream_t to launch on; NULL = legacy default stream */
  double  param;     /* the OtherPars double of setvaltobad (value) / setbadtoval (newval): $COMP(...) */
  /* set{nan,inf,nonfinite}tobad mark their output BAD only if they wrote a BAD value (`if (flag)
   * $PDLSTATESETBAD(b)`, lib/PDL/Bad.pd:695-707).  Those three ops REQUIRE a host pointer here; the
   * call synchronises the stream and stores 1/0 (unless tflags has PDLB200_TRANS_DEFER_ANYBAD).  minmaximum
   * uses it the same way for "a row had no usable element" (Ufunc.pd:578-583), and _n_ind for "a slot could not be
   * filled".  Ignored by every other op. */
  int32_t *anybad;
} pdlb200_trans;

/* Return codes.  Non-zero => `err` (if given) holds a NUL-terminated message the
 * shim turns into PDL->make_error(PDL_EUSERERROR|PDL_EFATAL, ...)  pdlutil.c:27-58 */
enum {
  PDLB200_OK = 0,
  PDLB200_EINVAL = 1,     /* malformed descriptor (user error) */
  PDLB200_EUNSUPPORTED = 2, /* op/type combination not on the device path */
  PDLB200_ENODEVICE = 3,  /* no CUDA device / driver */
  PDLB200_ECUDA = 4       /* CUDA runtime error */
};

/* --- the hot path ------------------------------------------------------- */
/* Replaces pdl_<op>_readdata for every PDLB200_OP_*; dispatches to the three
 * family launchers below. */
PDLB200_API int pdlb200_readdata(const pdlb200_trans *t, char *err, size_t errlen);
/* Ops.pd biop/bifunc/ufunc bodies on the broadcast loop. */
PDLB200_API int pdlb200_elementwise(const pdlb200_trans *t, char *err, size_t errlen);
/* Ufunc.pd a(n) reductions and scans. */
PDLB200_API int pdlb200_reduce(const pdlb200_trans *t, char *err, size_t errlen);
/* Primitive.pd matmult. */
PDLB200_API int pdlb200_matmult(const pdlb200_trans *t, char *err, size_t errlen);

/* --- device data store (north-star subsystem 1) ---------------------------
 * Replaces pdl_allocdata (pdlapi.c:172-209) / pdl__free data (pdlapi.c:283-316)
 * for device-resident ndarrays: a stream-ordered pool, no zero-fill (every
 * kernel overwrites its whole output), lazy host sync by dirty bits. */
typedef struct pdlb200_buf pdlb200_buf;
PDLB200_API int    pdlb200_buf_new(size_t nbytes, pdlb200_buf **out, char *err, size_t errlen);
PDLB200_API void   pdlb200_buf_free(pdlb200_buf *b);
PDLB200_API size_t pdlb200_buf_nbytes(const pdlb200_buf *b);
/* Device pointer, valid for the buffer's lifetime.  for_write marks the device
 * copy newer than the host's. */
PDLB200_API void  *pdlb200_buf_devptr(pdlb200_buf *b, int for_write);
/* Host -> device: upload `nbytes` from `host` (marks device copy current). */
PDLB200_API int    pdlb200_buf_upload(pdlb200_buf *b, const void *host, size_t nbytes, void *stream, char *err, size_t errlen);
/* Device -> host iff the device copy is newer (or force != 0); synchronises the stream. */
PDLB200_API int    pdlb200_buf_download(pdlb200_buf *b, void *host, size_t nbytes, int force, void *stream, char *err, size_t errlen);
PDLB200_API int    pdlb200_buf_device_dirty(const pdlb200_buf *b);
/* The allocator under the store, usable directly: device memory without zero-fill from an exact-size free list in
 * front of the stream-ordered pool (a recycled block costs a hash lookup).  `nbytes` of free must be the allocation's. */
PDLB200_API void  *pdlb200_dev_alloc(size_t nbytes);
PDLB200_API void   pdlb200_dev_free(void *p, size_t nbytes);
PDLB200_API void   pdlb200_dev_trim(void);            /* hand the cached blocks back to the pool */

/* --- plumbing ------------------------------------------------------------- */
PDLB200_API int    pdlb200_abi_version(void);
/* Identity of the sources this library was built from (pdl_b200/build.py source_id): the Python loader refuses a
 * library whose id does not match the sources next to it. */
PDLB200_API const char *pdlb200_build_id(void);
PDLB200_API int    pdlb200_device_count(void);                 /* 0 when no usable GPU */
PDLB200_API int    pdlb200_set_device(int dev, char *err, size_t errlen);
PDLB200_API int    pdlb200_sm_count(void);
PDLB200_API int    pdlb200_sync(void *stream, char *err, size_t errlen);
/* Pinned host memory for the e2e path (host<->device copies at PCIe rate). */
PDLB200_API void  *pdlb200_host_alloc(size_t nbytes);
/* Same, write-combined: for staging buffers the host only fills sequentially and the GPU reads (H2D), the
 * pages are not snooped, which helps when several GPUs pull from host memory at once. */
PDLB200_API void  *pdlb200_host_alloc_wc(size_t nbytes);
PDLB200_API void   pdlb200_host_free(void *p);
PDLB200_API int    pdlb200_memcpy_h2d(void *dst, const void *src, size_t nbytes, void *stream, char *err, size_t errlen);
PDLB200_API int    pdlb200_memcpy_d2h(void *dst, const void *src, size_t nbytes, void *stream, char *err, size_t errlen);
/* Unified (CUDA managed) memory: what the XS shim backs ndarray data with, so that the
 * UNMODIFIED reference core can keep dereferencing pdl->data on the host while kernels use the
 * same pointer; pages migrate on demand in both directions (= lazy host sync, done by the
 * driver), and chained device ops never cross PCIe.  See perl/PDL-B200/B200.xs. */
PDLB200_API void  *pdlb200_managed_alloc(size_t nbytes);
PDLB200_API void   pdlb200_managed_free(void *p);   /* returns the block to an exact-size free list */
PDLB200_API void   pdlb200_managed_trim(void);      /* hand every cached block back to the driver */
/* 0 = plain host memory, 1 = device memory, 2 = managed, 3 = pinned host (device-visible) */
PDLB200_API int    pdlb200_ptr_kind(const void *p);
/* Migrate a managed range towards the device (to_device != 0) or the host ahead of use. */
PDLB200_API int    pdlb200_prefetch(void *p, size_t nbytes, int to_device, void *stream, char *err, size_t errlen);
/* --- coherent host/device store for an UNMODIFIED host core (north-star subsystem 1) -------------------------
 * What the reference-side binding (perl/PDL-B200) backs ndarray data with.  One buffer = a cudaMalloc'd
 * (stream-ordered pool) device allocation + a page-aligned host MIRROR of the same size, which is what the host
 * core sees as pdl->data (replaces the SV of pdl_allocdata, lib/PDL/Core/pdlapi.c:172-209) + host/device dirty
 * bits.  While the mirror is not current its pages are PROT_NONE: the binding makes it current explicitly at the
 * reference's host-access choke points it can see (pdlb200_mbuf_host), and ANY other host dereference
 * (lib/PDL/Core.xs:771-859,1045,1145-1199, pdlconv.c:6-43, CPU readdata bodies) faults into the library's
 * SIGSEGV handler, which does the same — stream sync, one D2H copy of the buffer, mprotect — and resumes; a
 * host WRITE additionally marks the device copy stale (re-uploaded by the next device op that reads it).
 * Chained device ops never cross PCIe and never synchronise.  Freed buffers are recycled by exact size. */
/* New buffer (contents undefined, mirror not current).  Returns the MIRROR address (the handle), NULL on failure. */
PDLB200_API void  *pdlb200_mbuf_new(size_t nbytes);
/* New buffer whose device copy holds `nbytes` from plain host memory `src` (one H2D copy; `src` may be released
 * on return).  The mirror is not populated. */
PDLB200_API void  *pdlb200_mbuf_adopt(const void *src, size_t nbytes, char *err, size_t errlen);
PDLB200_API void   pdlb200_mbuf_retain(void *mirror);   /* one more owner (an aliasing view, e.g. clump of a physical parent) */
PDLB200_API void   pdlb200_mbuf_free(void *mirror);     /* drop one owner; the last one recycles the buffer */
PDLB200_API int    pdlb200_mbuf_is(const void *mirror); /* 1 iff `mirror` is the base address of a live store buffer */
/* Device pointer for a kernel launched on `stream`.  A stale device copy is uploaded first unless `discard`
 * (the kernel overwrites the whole buffer); `for_write` marks the mirror stale. */
PDLB200_API void  *pdlb200_mbuf_dev(void *mirror, int for_write, int discard, void *stream, char *err, size_t errlen);
/* Make the mirror current (download iff the device copy is newer); `for_write` marks the device copy stale.
 * No-op (OK) for pointers that are not store buffers. */
PDLB200_API int    pdlb200_mbuf_host(void *mirror, int for_write, char *err, size_t errlen);
/* -1 not a store buffer; else bit0-1 host state (0 stale, 1 current clean, 2 current + host-modified), bit2 device copy current */
PDLB200_API int    pdlb200_mbuf_state(const void *mirror);
/* out[8] = buffers created, recycled, uploads, upload bytes, downloads, download bytes, faults handled, adopted */
PDLB200_API void   pdlb200_mbuf_stats(uint64_t *out);
PDLB200_API void   pdlb200_mbuf_trim(void);            /* hand every cached buffer back to the driver / the OS */
/* Which transformation vtables of the host core run on the device: the binding registers them when it attaches
 * and asks from its PDL->make_trans_mutual wrapper (everything else is a CPU op whose parameters must be made
 * current in host memory first).  Opaque pointers. */
PDLB200_API void   pdlb200_devop_register(const void *vtable, int on);
PDLB200_API int    pdlb200_devop_is(const void *vtable);
/* --- record exchange over peer memory (NVLink / NVSwitch P2P) for the sharded reductions ---------------------------
 * Step 2 of a sharded whole-array reduction (PART_ -> gather -> COLL_) without a library collective: every rank
 * owns a mailbox in its own HBM that all peers map through CUDA IPC; pdlb200_peer_exchange() is ONE kernel that
 * stores this rank's records into every peer's mailbox over NVLink, publishes an epoch flag (release, system scope)
 * and waits for the flags of all peers.  The COLL_* launch that follows reads the gathered records locally. */
PDLB200_API void  *pdlb200_peer_mailbox_new(int world, size_t cap_words);   /* cudaMalloc'd, zeroed, IPC-exportable; cap = int64 words per rank */
PDLB200_API void   pdlb200_peer_mailbox_free(void *mailbox);
PDLB200_API int    pdlb200_ipc_export(void *devptr, unsigned char *handle64, char *err, size_t errlen);
PDLB200_API void  *pdlb200_ipc_open(const unsigned char *handle64, char *err, size_t errlen);
PDLB200_API void   pdlb200_ipc_close(void *mapped);
/* `mailboxes`: DEVICE array of `world` device pointers (rank order; this rank's own mailbox included).  `epoch` is
 * the same on all ranks and increases by 1 per exchange, starting at 1. */
PDLB200_API int    pdlb200_peer_exchange(const void *local, size_t nwords, size_t cap_words, void *const *mailboxes, int rank,
                                         int world, int64_t epoch, void *stream, char *err, size_t errlen);
/* int64-word offset of the gathered [world][cap_words] records of `epoch` inside a mailbox (rank r's at + r * cap_words) */
PDLB200_API int64_t pdlb200_peer_gathered_offset(int world, size_t cap_words, int64_t epoch);
/* Number of kernels this library has launched in this process (bench "gpu_launches"). */
PDLB200_API uint64_t pdlb200_launch_count(void);
/* Name of the kernel variant chosen by the most recent launch on this thread (introspection,
 * the analogue of get_autopthread_actual, lib/PDL/Core.xs:564-573). */
PDLB200_API const char *pdlb200_last_kernel(void);
PDLB200_API const char *pdlb200_op_name(int op);
PDLB200_API size_t pdlb200_type_size(int type);

#ifdef __cplusplus
}
#endif
#endif /* PDLB200_H */
