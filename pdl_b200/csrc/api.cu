// api.cu — the extern "C" surface of libpdlb200 (include/pdlb200.h): descriptor
// validation, family dispatch, the device data store and plumbing.  No torch, no
// Perl: plain pointers and sizes only.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include "common.cuh"

namespace pdlb200 {

static std::atomic<uint64_t> g_launches{0};
static thread_local const char *g_last_kernel = "";
static int g_sm_count = 0;
static int g_dev_count = -1;
static std::mutex g_mu;

void note_launch(const char *name) { g_launches.fetch_add(1, std::memory_order_relaxed); g_last_kernel = name; }

static int probe_devices() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_dev_count >= 0) return g_dev_count;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
  g_dev_count = n;
  return n;
}

int sm_count() {
  if (g_sm_count > 0) return g_sm_count;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) == cudaSuccess &&
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
    g_sm_count = n;
  else
    g_sm_count = 148;  // B200
  return g_sm_count;
}

// scratch for two-stage reductions / scans / flag words: one grow-only buffer per (device, stream), used in that
// stream's order — two streams running scratch-using ops at the same time never share a buffer.
struct ScratchBuf { void *p; size_t sz; };
static std::unordered_map<uint64_t, ScratchBuf> g_scratch;
void *scratch(size_t nbytes, cudaStream_t s) {
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t key = ((uint64_t)(uintptr_t)s << 4) ^ (uint64_t)(dev & 15);
  std::lock_guard<std::mutex> lk(g_mu);
  ScratchBuf &b = g_scratch[key];
  if (b.sz < nbytes) {
    if (b.p) { cudaStreamSynchronize(s); cudaFree(b.p); }
    const size_t want = nbytes < (1u << 20) ? (1u << 20) : nbytes;
    if (cudaMalloc(&b.p, want) != cudaSuccess) { cudaGetLastError(); b.p = nullptr; b.sz = 0; return nullptr; }
    b.sz = want;
  }
  return b.p;
}

static std::unordered_multimap<size_t, void *> g_managed_free;   // exact-size free list
static std::unordered_map<void *, size_t> g_managed_size;         // live managed blocks
static size_t g_managed_cached = 0;
static size_t managed_cache_cap() {
  static size_t cap = 0;
  if (!cap) {
    const char *e = getenv("PDLB200_MANAGED_CACHE_MB");
    cap = (size_t)(e ? atoll(e) : 16384) << 20;
    if (!cap) cap = 1;
  }
  return cap;
}

static const char *op_names[PDLB200_OP__END] = {};
static void init_names() {
  static bool done = false;
  if (done) return;
  done = true;
#define N(ID, S) op_names[ID] = S;
  N(PDLB200_OP_PLUS, "plus") N(PDLB200_OP_MULT, "mult") N(PDLB200_OP_MINUS, "minus") N(PDLB200_OP_DIVIDE, "divide")
  N(PDLB200_OP_GT, "gt") N(PDLB200_OP_LT, "lt") N(PDLB200_OP_LE, "le") N(PDLB200_OP_GE, "ge")
  N(PDLB200_OP_EQ, "eq") N(PDLB200_OP_NE, "ne") N(PDLB200_OP_SHIFTLEFT, "shiftleft")
  N(PDLB200_OP_SHIFTRIGHT, "shiftright") N(PDLB200_OP_OR2, "or2") N(PDLB200_OP_AND2, "and2") N(PDLB200_OP_XOR, "xor")
  N(PDLB200_OP_POWER, "power") N(PDLB200_OP_ATAN2, "atan2") N(PDLB200_OP_MODULO, "modulo") N(PDLB200_OP_SPACESHIP, "spaceship")
  N(PDLB200_OP_BITNOT, "bitnot") N(PDLB200_OP_SQRT, "sqrt") N(PDLB200_OP_SIN, "sin") N(PDLB200_OP_COS, "cos")
  N(PDLB200_OP_NOT, "not") N(PDLB200_OP_EXP, "exp") N(PDLB200_OP_LOG, "log") N(PDLB200_OP_LOG10, "log10")
  N(PDLB200_OP_RABS, "_rabs") N(PDLB200_OP_ASSGN, "assgn") N(PDLB200_OP_ABS2, "abs2")
  N(PDLB200_OP_SUMOVER, "sumover") N(PDLB200_OP_PRODOVER, "prodover") N(PDLB200_OP_DSUMOVER, "dsumover")
  N(PDLB200_OP_DPRODOVER, "dprodover") N(PDLB200_OP_AVERAGE, "average") N(PDLB200_OP_DAVERAGE, "daverage")
  N(PDLB200_OP_MINIMUM, "minimum") N(PDLB200_OP_MAXIMUM, "maximum") N(PDLB200_OP_MINIMUM_IND, "minimum_ind")
  N(PDLB200_OP_MAXIMUM_IND, "maximum_ind") N(PDLB200_OP_ANDOVER, "andover") N(PDLB200_OP_OROVER, "orover")
  N(PDLB200_OP_BANDOVER, "bandover") N(PDLB200_OP_BOROVER, "borover") N(PDLB200_OP_ZCOVER, "zcover")
  N(PDLB200_OP_XOROVER, "xorover") N(PDLB200_OP_BXOROVER, "bxorover")
  N(PDLB200_OP_NBADOVER, "nbadover") N(PDLB200_OP_NGOODOVER, "ngoodover")
  N(PDLB200_OP_CUMUSUMOVER, "cumusumover") N(PDLB200_OP_CUMUPRODOVER, "cumuprodover")
  N(PDLB200_OP_DCUMUSUMOVER, "dcumusumover") N(PDLB200_OP_DCUMUPRODOVER, "dcumuprodover")
  N(PDLB200_OP_MATMULT, "matmult") N(PDLB200_OP_CONVERT, "converttype") N(PDLB200_OP_IPOW, "ipow")
  N(PDLB200_OP_ISBAD, "isbad") N(PDLB200_OP_ISGOOD, "isgood") N(PDLB200_OP_ISNAN, "isnan")
  N(PDLB200_OP_SETBADIF, "setbadif") N(PDLB200_OP_SETVALTOBAD, "setvaltobad") N(PDLB200_OP_SETNANTOBAD, "setnantobad")
  N(PDLB200_OP_SETINFTOBAD, "setinftobad") N(PDLB200_OP_SETNONFINITETOBAD, "setnonfinitetobad")
  N(PDLB200_OP_SETBADTONAN, "setbadtonan") N(PDLB200_OP_SETBADTOVAL, "setbadtoval") N(PDLB200_OP_BADMASK, "badmask")
  N(PDLB200_OP_COPYBAD, "copybad") N(PDLB200_OP_AXISVALUES, "axisvalues") N(PDLB200_OP_INNER, "inner")
  N(PDLB200_OP_MINMAXIMUM, "minmaximum") N(PDLB200_OP_MAGNOVER, "magnover") N(PDLB200_OP_OUTER, "outer")
  N(PDLB200_OP_PART_SUM, "part_sum") N(PDLB200_OP_PART_DSUM, "part_dsum") N(PDLB200_OP_PART_MIN, "part_min")
  N(PDLB200_OP_PART_MAX, "part_max") N(PDLB200_OP_COLL_SUM, "coll_sum") N(PDLB200_OP_COLL_AVG, "coll_avg")
  N(PDLB200_OP_COLL_MIN, "coll_min") N(PDLB200_OP_COLL_MAX, "coll_max") N(PDLB200_OP_COLL_MIN_IND, "coll_min_ind")
  N(PDLB200_OP_COLL_MAX_IND, "coll_max_ind")
  N(PDLB200_OP_MINIMUM_N_IND, "minimum_n_ind") N(PDLB200_OP_MAXIMUM_N_IND, "maximum_n_ind")
#undef N
}

static int validate(const pdlb200_trans *t, const Err &E) {
  if (!t) return E.fail(PDLB200_EINVAL, "pdlb200: NULL descriptor");
  if (t->op < 0 || t->op >= PDLB200_OP__END || !pdlb200_op_name(t->op)[0])
    return E.fail(PDLB200_EINVAL, "pdlb200: unknown op %d", t->op);
  if (t->datatype < 0) return E.fail(PDLB200_EINVAL, "%s: invalid datatype %d", pdlb200_op_name(t->op), t->datatype);
  const bool cplx = (t->datatype == PDLB200_CF || t->datatype == PDLB200_CD) && t->op >= PDLB200_OP_PLUS && t->op <= PDLB200_OP_DIVIDE;
  if (t->datatype >= PDLB200_NTYPES && !cplx)
    return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d (long double / complex) has no device representation",
                  pdlb200_op_name(t->op), t->datatype);
  if (t->ndims < 0 || t->ndims > PDLB200_MAXDIMS)
    return E.fail(PDLB200_EINVAL, "%s: ndims %d out of range", pdlb200_op_name(t->op), t->ndims);
  if (t->npdls < 2 || t->npdls > PDLB200_MAXPDLS)
    return E.fail(PDLB200_EINVAL, "%s: npdls %d out of range", pdlb200_op_name(t->op), t->npdls);
  for (int d = 0; d < t->ndims; d++)
    if (t->dims[d] < 0) return E.fail(PDLB200_EINVAL, "%s: broadcast dim %d has size %lld", pdlb200_op_name(t->op), d, (long long)t->dims[d]);
  if (probe_devices() <= 0)
    return E.fail(PDLB200_ENODEVICE, "%s: no CUDA device available and libpdlb200 has no CPU fallback", pdlb200_op_name(t->op));
  return PDLB200_OK;
}

int ew_arith(const pdlb200_trans *, const Err &);
int ew_cmp(const pdlb200_trans *, const Err &);
int ew_bits(const pdlb200_trans *, const Err &);
int ew_func(const pdlb200_trans *, const Err &);
int ew_unary(const pdlb200_trans *, const Err &);

int launch_elementwise(const pdlb200_trans *t, const Err &E) {
  const int op = t->op;
  if (t->datatype == PDLB200_CF || t->datatype == PDLB200_CD) return launch_complex(t, E);
  if (op <= PDLB200_OP_DIVIDE || op == PDLB200_OP_OUTER) return ew_arith(t, E);
  if (op <= PDLB200_OP_NE) return ew_cmp(t, E);
  if (op <= PDLB200_OP_XOR || op == PDLB200_OP_BITNOT) return ew_bits(t, E);
  if (op <= PDLB200_OP_SPACESHIP) return ew_func(t, E);
  if (op <= PDLB200_OP_ABS2) return ew_unary(t, E);
  if (op == PDLB200_OP_CONVERT) return launch_convert(t, E);
  if (op == PDLB200_OP_IPOW) return launch_ipow(t, E);
  if (op >= PDLB200_OP_ISBAD && op <= PDLB200_OP_COPYBAD) return launch_badops(t, E);
  if (op == PDLB200_OP_AXISVALUES) return launch_axisvalues(t, E);
  return E.fail(PDLB200_EINVAL, "%s is not an elementwise op", pdlb200_op_name(op));
}

}  // namespace pdlb200

using namespace pdlb200;

extern "C" {

int pdlb200_abi_version(void) { return PDLB200_ABI_VERSION; }
#ifndef PDLB200_BUILD_ID
#define PDLB200_BUILD_ID "unknown"
#endif
const char *pdlb200_build_id(void) { return PDLB200_BUILD_ID; }
int pdlb200_device_count(void) { return probe_devices(); }
int pdlb200_sm_count(void) { return probe_devices() > 0 ? sm_count() : 0; }
uint64_t pdlb200_launch_count(void) { return g_launches.load(); }
const char *pdlb200_last_kernel(void) { return g_last_kernel; }
const char *pdlb200_op_name(int op) {
  init_names();
  if (op < 0 || op >= PDLB200_OP__END || !op_names[op]) return "";
  return op_names[op];
}
size_t pdlb200_type_size(int type) {
  static const size_t sz[15] = {1, 1, 2, 2, 4, 4, 8, 8, 8, 4, 8, 0, 8, 16, 0};
  return (type >= 0 && type < 15) ? sz[type] : 0;
}

int pdlb200_set_device(int dev, char *err, size_t errlen) {
  Err E{err, errlen};
  if (probe_devices() <= 0) return E.fail(PDLB200_ENODEVICE, "pdlb200_set_device: no CUDA device available");
  PDLB200_CUDA_OK(cudaSetDevice(dev), E);
  g_sm_count = 0;
  return PDLB200_OK;
}

int pdlb200_sync(void *stream, char *err, size_t errlen) {
  Err E{err, errlen};
  if (probe_devices() <= 0) return E.fail(PDLB200_ENODEVICE, "pdlb200_sync: no CUDA device available");
  PDLB200_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream), E);
  return PDLB200_OK;
}

int pdlb200_elementwise(const pdlb200_trans *t, char *err, size_t errlen) {
  Err E{err, errlen};
  if (int rc = validate(t, E)) return rc;
  return launch_elementwise(t, E);
}
int pdlb200_reduce(const pdlb200_trans *t, char *err, size_t errlen) {
  Err E{err, errlen};
  if (int rc = validate(t, E)) return rc;
  if (t->op >= PDLB200_OP_CUMUSUMOVER && t->op <= PDLB200_OP_DCUMUPRODOVER) return launch_scan(t, E);
  if (t->op >= PDLB200_OP_SUMOVER && t->op <= PDLB200_OP_NGOODOVER) return launch_reduce(t, E);
  if (t->op == PDLB200_OP_INNER) return launch_inner(t, E);
  if (t->op == PDLB200_OP_MINMAXIMUM) return launch_minmaximum(t, E);
  if (t->op == PDLB200_OP_MAGNOVER) return launch_magnover(t, E);
  if (t->op >= PDLB200_OP_PART_SUM && t->op <= PDLB200_OP_PART_MAX) return launch_partial(t, E);
  if (t->op >= PDLB200_OP_COLL_SUM && t->op <= PDLB200_OP_COLL_MAX_IND) return launch_collapse(t, E);
  if (t->op == PDLB200_OP_MINIMUM_N_IND || t->op == PDLB200_OP_MAXIMUM_N_IND) return launch_nind(t, E);
  return E.fail(PDLB200_EINVAL, "%s is not a reduction", pdlb200_op_name(t->op));
}
int pdlb200_matmult(const pdlb200_trans *t, char *err, size_t errlen) {
  Err E{err, errlen};
  if (int rc = validate(t, E)) return rc;
  if (t->op != PDLB200_OP_MATMULT) return E.fail(PDLB200_EINVAL, "%s is not matmult", pdlb200_op_name(t->op));
  return launch_matmult(t, E);
}
int pdlb200_readdata(const pdlb200_trans *t, char *err, size_t errlen) {
  Err E{err, errlen};
  if (int rc = validate(t, E)) return rc;
  const int op = t->op;
  if (op <= PDLB200_OP_ABS2 || op == PDLB200_OP_CONVERT || op == PDLB200_OP_IPOW ||
      (op >= PDLB200_OP_ISBAD && op <= PDLB200_OP_AXISVALUES) || op == PDLB200_OP_OUTER) return launch_elementwise(t, E);
  if (op == PDLB200_OP_INNER) return launch_inner(t, E);
  if (op == PDLB200_OP_MINMAXIMUM) return launch_minmaximum(t, E);
  if (op == PDLB200_OP_MAGNOVER) return launch_magnover(t, E);
  if (op >= PDLB200_OP_CUMUSUMOVER && op <= PDLB200_OP_DCUMUPRODOVER) return launch_scan(t, E);
  if (op >= PDLB200_OP_SUMOVER && op <= PDLB200_OP_NGOODOVER) return launch_reduce(t, E);
  if (op == PDLB200_OP_MATMULT) return launch_matmult(t, E);
  if (op >= PDLB200_OP_PART_SUM && op <= PDLB200_OP_PART_MAX) return launch_partial(t, E);
  if (op >= PDLB200_OP_COLL_SUM && op <= PDLB200_OP_COLL_MAX_IND) return launch_collapse(t, E);
  if (op == PDLB200_OP_MINIMUM_N_IND || op == PDLB200_OP_MAXIMUM_N_IND) return launch_nind(t, E);
  return E.fail(PDLB200_EINVAL, "pdlb200: op %d has no launcher", op);
}

// ---- device data store ---------------------------------------------------------
struct pdlb200_buf {
  void *dev;
  size_t nbytes;
  int device;
  int dev_dirty;  // device copy newer than any host copy
};

// Raw device allocations behind the store: an exact-size free list in front of the stream-ordered pool.  PDL code
// creates same-sized temporaries over and over (`$x = $y + $c` makes a fresh 32 MiB output per op): a recycled block
// costs a hash lookup instead of a cudaMallocAsync + cudaFreeAsync pair (~10 us against a 17 us kernel), and there is
// no zero-fill (the reference's memset in pdl_allocdata, pdlapi.c:204, is ~75% of its config-1 time).  Reuse is
// ordered by the launch stream, like the pool itself.
static std::unordered_multimap<size_t, void *> g_dev_free[16];
static size_t g_dev_cached[16] = {0};
static size_t dev_cache_cap() {
  static size_t cap = 0;
  if (!cap) { const char *e = getenv("PDLB200_DEV_CACHE_MB"); cap = (size_t)(e ? atoll(e) : 65536) << 20; if (!cap) cap = 1; }
  return cap;
}
void *pdlb200_dev_alloc(size_t nbytes) {
  if (probe_devices() <= 0) return nullptr;
  if (!nbytes) nbytes = 1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_dev_free[dev].find(nbytes);
    if (it != g_dev_free[dev].end()) {
      void *p = it->second;
      g_dev_free[dev].erase(it);
      g_dev_cached[dev] -= nbytes;
      return p;
    }
  }
  static bool pool_ready[16] = {false};
  if (!pool_ready[dev]) {
    // keep freed blocks in the pool instead of returning them to the driver at every sync
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    pool_ready[dev] = true;
  }
  void *p = nullptr;
  if (cudaMallocAsync(&p, nbytes, (cudaStream_t)0) != cudaSuccess) {
    cudaGetLastError();
    pdlb200_dev_trim();
    if (cudaMallocAsync(&p, nbytes, (cudaStream_t)0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  }
  return p;
}
void pdlb200_dev_free(void *p, size_t nbytes) {
  if (!p) return;
  if (!nbytes) nbytes = 1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 16) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_dev_cached[dev] + nbytes <= dev_cache_cap()) { g_dev_free[dev].emplace(nbytes, p); g_dev_cached[dev] += nbytes; return; }
  }
  cudaFreeAsync(p, (cudaStream_t)0);
}
void pdlb200_dev_trim(void) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return;
  std::unordered_multimap<size_t, void *> drop;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    drop.swap(g_dev_free[dev]);
    g_dev_cached[dev] = 0;
  }
  for (auto &kv : drop) cudaFreeAsync(kv.second, (cudaStream_t)0);
  cudaStreamSynchronize((cudaStream_t)0);
}

int pdlb200_buf_new(size_t nbytes, pdlb200_buf **out, char *err, size_t errlen) {
  Err E{err, errlen};
  if (!out) return E.fail(PDLB200_EINVAL, "pdlb200_buf_new: NULL out");
  if (probe_devices() <= 0) return E.fail(PDLB200_ENODEVICE, "pdlb200_buf_new: no CUDA device available");
  pdlb200_buf *b = new pdlb200_buf{nullptr, nbytes, 0, 0};
  cudaGetDevice(&b->device);
  if (nbytes) {
    b->dev = pdlb200_dev_alloc(nbytes);
    if (!b->dev) { delete b; return E.fail(PDLB200_ECUDA, "pdlb200_buf_new: cannot allocate %zu bytes on the device", nbytes); }
  }
  *out = b;
  return PDLB200_OK;
}
void pdlb200_buf_free(pdlb200_buf *b) {
  if (!b) return;
  if (b->dev) pdlb200_dev_free(b->dev, b->nbytes);
  delete b;
}
size_t pdlb200_buf_nbytes(const pdlb200_buf *b) { return b ? b->nbytes : 0; }
void *pdlb200_buf_devptr(pdlb200_buf *b, int for_write) {
  if (!b) return nullptr;
  if (for_write) b->dev_dirty = 1;
  return b->dev;
}
int pdlb200_buf_device_dirty(const pdlb200_buf *b) { return b ? b->dev_dirty : 0; }
int pdlb200_buf_upload(pdlb200_buf *b, const void *host, size_t nbytes, void *stream, char *err, size_t errlen) {
  Err E{err, errlen};
  if (!b || nbytes > b->nbytes) return E.fail(PDLB200_EINVAL, "pdlb200_buf_upload: bad buffer or size");
  if (nbytes) PDLB200_CUDA_OK(cudaMemcpyAsync(b->dev, host, nbytes, cudaMemcpyHostToDevice, (cudaStream_t)stream), E);
  b->dev_dirty = 0;
  return PDLB200_OK;
}
int pdlb200_buf_download(pdlb200_buf *b, void *host, size_t nbytes, int force, void *stream, char *err, size_t errlen) {
  Err E{err, errlen};
  if (!b || nbytes > b->nbytes) return E.fail(PDLB200_EINVAL, "pdlb200_buf_download: bad buffer or size");
  if (!b->dev_dirty && !force) return PDLB200_OK;
  if (nbytes) PDLB200_CUDA_OK(cudaMemcpyAsync(host, b->dev, nbytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream), E);
  PDLB200_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream), E);
  b->dev_dirty = 0;
  return PDLB200_OK;
}

// Managed buffers are recycled through an exact-size free list: PDL scripts create same-sized
// temporaries over and over (`$x = $y + $c` makes a fresh 32 MiB output per op), and
// cudaMallocManaged + first-touch population + cudaFree cost ~2 ms per op against a 25 us kernel.
// A recycled buffer is already resident in HBM; a new one is prefetched there so the producing
// kernel does not take GPU page faults.  Cap: PDLB200_MANAGED_CACHE_MB (default 16384).
void *pdlb200_managed_alloc(size_t nbytes) {
  if (probe_devices() <= 0) return nullptr;
  if (!nbytes) nbytes = 1;
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_managed_free.find(nbytes);
    if (it != g_managed_free.end()) {
      void *p = it->second;
      g_managed_free.erase(it);
      g_managed_cached -= nbytes;
      g_managed_size[p] = nbytes;
      return p;
    }
  }
  void *p = nullptr;
  if (cudaMallocManaged(&p, nbytes, cudaMemAttachGlobal) != cudaSuccess) {
    cudaGetLastError();
    pdlb200_managed_trim();   // give cached blocks back to the driver and retry once
    if (cudaMallocManaged(&p, nbytes, cudaMemAttachGlobal) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  }
  if (nbytes >= (1u << 16)) cudaMemPrefetchAsync(p, nbytes, dev, (cudaStream_t)0);
  std::lock_guard<std::mutex> lk(g_mu);
  g_managed_size[p] = nbytes;
  return p;
}
void pdlb200_managed_free(void *p) {
  if (!p) return;
  size_t nbytes = 0;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_managed_size.find(p);
    if (it != g_managed_size.end()) { nbytes = it->second; g_managed_size.erase(it); }
    if (nbytes && g_managed_cached + nbytes <= managed_cache_cap()) {
      g_managed_free.emplace(nbytes, p);
      g_managed_cached += nbytes;
      return;
    }
  }
  cudaFree(p);
}
void pdlb200_managed_trim(void) {
  std::unordered_multimap<size_t, void *> drop;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    drop.swap(g_managed_free);
    g_managed_cached = 0;
  }
  for (auto &kv : drop) cudaFree(kv.second);
}
int pdlb200_ptr_kind(const void *p) {
  if (!p || probe_devices() <= 0) return 0;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return 0; }
  switch (a.type) {
    case cudaMemoryTypeDevice: return 1;
    case cudaMemoryTypeManaged: return 2;
    case cudaMemoryTypeHost: return 3;
    default: return 0;
  }
}
int pdlb200_prefetch(void *p, size_t nbytes, int to_device, void *stream, char *err, size_t errlen) {
  Err E{err, errlen};
  if (probe_devices() <= 0) return E.fail(PDLB200_ENODEVICE, "pdlb200_prefetch: no CUDA device available");
  int dev = 0;
  PDLB200_CUDA_OK(cudaGetDevice(&dev), E);
  PDLB200_CUDA_OK(cudaMemPrefetchAsync(p, nbytes, to_device ? dev : cudaCpuDeviceId, (cudaStream_t)stream), E);
  return PDLB200_OK;
}

void *pdlb200_host_alloc(size_t nbytes) {
  if (probe_devices() <= 0) return nullptr;
  void *p = nullptr;
  if (cudaMallocHost(&p, nbytes ? nbytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void *pdlb200_host_alloc_wc(size_t nbytes) {
  if (probe_devices() <= 0) return nullptr;
  void *p = nullptr;
  if (cudaHostAlloc(&p, nbytes ? nbytes : 1, cudaHostAllocWriteCombined) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void pdlb200_host_free(void *p) { if (p) cudaFreeHost(p); }
int pdlb200_memcpy_h2d(void *dst, const void *src, size_t nbytes, void *stream, char *err, size_t errlen) {
  Err E{err, errlen};
  if (probe_devices() <= 0) return E.fail(PDLB200_ENODEVICE, "pdlb200_memcpy_h2d: no CUDA device available");
  PDLB200_CUDA_OK(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, (cudaStream_t)stream), E);
  return PDLB200_OK;
}
int pdlb200_memcpy_d2h(void *dst, const void *src, size_t nbytes, void *stream, char *err, size_t errlen) {
  Err E{err, errlen};
  if (probe_devices() <= 0) return E.fail(PDLB200_ENODEVICE, "pdlb200_memcpy_d2h: no CUDA device available");
  PDLB200_CUDA_OK(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream), E);
  return PDLB200_OK;
}

}  // extern "C"
