// store.cu — the device-resident ndarray data store for an UNMODIFIED host core (north-star subsystem 1).
//
// Replaces pdl_allocdata / the data half of pdl__free (lib/PDL/Core/pdlapi.c:172-209,283-316) for ndarrays the
// device path creates or adopts.  One buffer ("mbuf") = a cudaMalloc'd (stream-ordered pool) device allocation
// + a host MIRROR of the same size that the host core sees as pdl->data, + host/device dirty bits:
//
//   host state  H0  mirror not current: the pages are PROT_NONE (and usually not even populated)
//               HR  mirror current and clean:  PROT_READ
//               HW  mirror current, possibly modified by the host: PROT_READ|PROT_WRITE, device copy stale
//   dev_valid       the device copy is current
//
// Device ops ask for the device pointer (mbuf_dev): a stale device copy is uploaded first, a written one marks
// the mirror H0.  Chained device ops therefore never cross PCIe and never synchronise.  Host code reaches the
// data through the reference's few choke points (lib/PDL/Core.xs:771-859,1045,1145-1199 at/listref/sclr/
// get_dataref, pdlconv.c:6-43 readdata/writebackdata_vaffine, the make_physical loop of every CPU op
// pdlapi.c:102-110): the binding calls mbuf_host() at the ones it can see (the Core function table), and every
// other dereference of a non-current mirror lands in the SIGSEGV handler below, which does the same thing —
// stream sync, ONE cudaMemcpy D2H of the whole buffer, mprotect — and resumes.  A read leaves both copies valid
// (HR); a write (x86 page-fault error code bit 1) makes the host copy the only valid one (HW).
//
// Freed buffers go to an exact-size free list (PDL scripts create same-sized temporaries over and over):
// steady state has no cudaMalloc/mmap per op.
#include <atomic>
#include <cerrno>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>
#include <fcntl.h>
#include <sys/mman.h>
#include <ucontext.h>
#include <unistd.h>
#include "common.cuh"

// The four device-memory primitives the store uses.  PDLB200_STORE_HOSTSIM exists for ONE purpose: tests/ compile this
// file a second time into a test-only library in which "device" memory is plain malloc'd host memory, so that the
// dirty-bit / mprotect / fault-handler state machine can be unit-tested on a box without a GPU
// (tests/test_store_sim.py).  libpdlb200.so is never built with it.
#ifdef PDLB200_STORE_HOSTSIM
static int dev_alloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 1; }
static void dev_free(void *p) { free(p); }
static int dev_h2d(void *d, const void *h, size_t n, void *) { memcpy(d, h, n); return 0; }
static int dev_d2h(void *h, const void *d, size_t n, int) { memcpy(h, d, n); return 0; }
static void dev_sync() {}
static int dev_current() { return 0; }
static void dev_pool_setup() {}
static const char *dev_err() { return "host simulation"; }
static int store_device_count() { return 1; }
namespace pdlb200 { int Err::fail(int code, const char *fmt, ...) const { if (buf && len) snprintf(buf, len, "%s", fmt); return code; } }
#else
static int dev_alloc(void **p, size_t n) { if (cudaMallocAsync(p, n, (cudaStream_t)0) == cudaSuccess) return 0; cudaGetLastError(); return 1; }
static void dev_free(void *p) { cudaFreeAsync(p, (cudaStream_t)0); }   // store buffers are recycled as (mirror, device) PAIRS by the store itself
static int dev_h2d(void *d, const void *h, size_t n, void *stream) {
  // pageable source: the call returns once the source has been consumed, so the caller may release / protect it
  return cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, (cudaStream_t)stream) == cudaSuccess ? 0 : 1;
}
static int dev_d2h(void *h, const void *d, size_t n, int device) {
  int cur = 0; cudaGetDevice(&cur);
  if (cur != device) cudaSetDevice(device);
  // legacy default stream: ordered after every kernel the binding launched on it
  const cudaError_t e = cudaMemcpy(h, d, n, cudaMemcpyDeviceToHost);
  if (cur != device) cudaSetDevice(cur);
  return e == cudaSuccess ? 0 : 1;
}
static void dev_sync() { cudaStreamSynchronize((cudaStream_t)0); }
static int dev_current() { int d = 0; cudaGetDevice(&d); return d; }
static void dev_pool_setup() {
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev_current()) == cudaSuccess) { uint64_t thr = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr); }
}
static const char *dev_err() { return cudaGetErrorString(cudaGetLastError()); }
static int store_device_count() { return pdlb200_device_count(); }
#endif

namespace pdlb200 {
namespace {

enum HostState { H0 = 0, HR = 1, HW = 2 };

// The mirror's pages are mapped TWICE from one memfd: `host` is the public mapping the host core sees (its
// protection carries the host state), `alias` is the store's own always-writable window on the same pages.
// Downloads land through `alias` while `host` is still PROT_NONE, so a second host thread (autopthread workers)
// that touches the buffer during the copy faults and waits instead of reading half-arrived data.
struct MBuf {
  char *host, *alias; void *dev; size_t nbytes, span;   // span = nbytes rounded up to pages (the mmap length)
  off_t off;                                            // offset of the pages in the arena file
  int hstate; int dev_valid; int device; int refs;
};

std::recursive_mutex g_mu;                       // recursive: a fault inside our own memcpy must not self-deadlock
std::map<uintptr_t, MBuf *> g_live;              // by mirror base address (ordered: the fault handler looks up ranges)
std::unordered_multimap<size_t, MBuf *> g_free;  // exact-size free list
size_t g_cached = 0;
std::atomic<uint64_t> g_stat[8];                 // 0 new 1 recycled 2 uploads 3 upload bytes 4 downloads 5 download bytes 6 faults 7 adopted
struct sigaction g_old_segv;
bool g_handler = false;
size_t g_page = 4096;
int g_arena_fd = -1;                              // memfd holding every mirror's pages
off_t g_arena_end = 0;

size_t cache_cap() {
  static size_t cap = 0;
  if (!cap) { const char *e = getenv("PDLB200_STORE_CACHE_MB"); cap = (size_t)(e ? atoll(e) : 32768) << 20; if (!cap) cap = 1; }
  return cap;
}

MBuf *find_base(const void *p) {
  auto it = g_live.find((uintptr_t)p);
  return it == g_live.end() ? nullptr : it->second;
}
MBuf *find_range(const void *p) {
  auto it = g_live.upper_bound((uintptr_t)p);
  if (it == g_live.begin()) return nullptr;
  --it;
  MBuf *b = it->second;
  return ((uintptr_t)p < (uintptr_t)b->host + b->span) ? b : nullptr;
}

void protect(MBuf *b, int st) {
  if (b->hstate == st) return;
  mprotect(b->host, b->span, st == H0 ? PROT_NONE : st == HR ? PROT_READ : (PROT_READ | PROT_WRITE));
  b->hstate = st;
}

// make the mirror current; for_write -> the host copy becomes the only valid one
int to_host(MBuf *b, bool for_write) {
  if (b->hstate == H0) {
    if (b->dev_valid && b->nbytes) {
      if (dev_d2h(b->alias, b->dev, b->nbytes, b->device) != 0) return -1;   // public mapping still PROT_NONE
      g_stat[4]++; g_stat[5] += b->nbytes;
    }
    protect(b, for_write ? HW : HR);
  }
  if (for_write) { protect(b, HW); b->dev_valid = 0; }
  return 0;
}

void segv_handler(int sig, siginfo_t *si, void *uctx) {
  MBuf *b = nullptr;
  bool ok = false;
  if (si && si->si_addr) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    b = find_range(si->si_addr);
    if (b) {
      bool is_write = true;
#if defined(__x86_64__)
      is_write = (((ucontext_t *)uctx)->uc_mcontext.gregs[REG_ERR] & 2) != 0;
#endif
      g_stat[6]++;
      ok = to_host(b, is_write) == 0;
    }
  }
  if (ok) return;                                // retry the faulting instruction
  // not ours (or the download failed): hand over to whoever was there before, else die the default way
  if (g_old_segv.sa_flags & SA_SIGINFO) { if (g_old_segv.sa_sigaction) { g_old_segv.sa_sigaction(sig, si, uctx); return; } }
  else if (g_old_segv.sa_handler != SIG_DFL && g_old_segv.sa_handler != SIG_IGN) { g_old_segv.sa_handler(sig); return; }
  signal(SIGSEGV, SIG_DFL);
}

void install_handler() {
  if (g_handler) return;
  g_page = (size_t)sysconf(_SC_PAGESIZE);
  struct sigaction sa;
  memset(&sa, 0, sizeof sa);
  sa.sa_sigaction = segv_handler;
  sa.sa_flags = SA_SIGINFO | SA_NODEFER;         // NODEFER: a nested fault (never expected) must not wedge the process
  sigemptyset(&sa.sa_mask);
  sigaction(SIGSEGV, &sa, &g_old_segv);
  g_handler = true;
}

void destroy_buf(MBuf *d) {
  if (d->dev) dev_free(d->dev);
  munmap(d->host, d->span); munmap(d->alias, d->span);
  fallocate(g_arena_fd, FALLOC_FL_PUNCH_HOLE | FALLOC_FL_KEEP_SIZE, d->off, (off_t)d->span);   // give the pages back
  delete d;
}

MBuf *alloc_buf(size_t nbytes) {
  {
    auto it = g_free.find(nbytes);
    if (it != g_free.end()) {
      MBuf *b = it->second;
      g_free.erase(it);
      g_cached -= b->span;
      g_stat[1]++;
      return b;
    }
  }
  install_handler();
  MBuf *b = new MBuf();
  b->nbytes = nbytes;
  b->span = (nbytes + g_page - 1) / g_page * g_page;
  if (!b->span) b->span = g_page;
  b->device = dev_current();
  if (g_arena_fd < 0) g_arena_fd = memfd_create("pdlb200-mirrors", MFD_CLOEXEC);
  if (g_arena_fd < 0) { delete b; return nullptr; }
  // 2 MiB-aligned offsets so that huge pages can back large mirrors where shmem THP is enabled
  const off_t align = b->span >= (2u << 20) ? (off_t)(2u << 20) : (off_t)g_page;
  b->off = (g_arena_end + align - 1) / align * align;
  if (ftruncate(g_arena_fd, b->off + (off_t)b->span) != 0) { delete b; return nullptr; }
  void *h = mmap(nullptr, b->span, PROT_NONE, MAP_SHARED | MAP_NORESERVE, g_arena_fd, b->off);
  void *al = h == MAP_FAILED ? MAP_FAILED : mmap(nullptr, b->span, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_NORESERVE, g_arena_fd, b->off);
  if (h == MAP_FAILED || al == MAP_FAILED) { if (h != MAP_FAILED) munmap(h, b->span); delete b; return nullptr; }
  g_arena_end = b->off + (off_t)b->span;
  b->host = (char *)h; b->alias = (char *)al;
  madvise(al, b->span, MADV_HUGEPAGE);
  if (dev_alloc(&b->dev, b->span) != 0) {
    // hand the cached pairs back and retry once
    std::vector<MBuf *> drop;
    for (auto &kv : g_free) drop.push_back(kv.second);
    g_free.clear(); g_cached = 0;
    for (MBuf *d : drop) destroy_buf(d);
    dev_sync();
    if (dev_alloc(&b->dev, b->span) != 0) { b->dev = nullptr; destroy_buf(b); return nullptr; }
  }
  g_stat[0]++;
  return b;
}

}  // namespace
}  // namespace pdlb200

using namespace pdlb200;

extern "C" {

void *pdlb200_mbuf_new(size_t nbytes) {
  if (store_device_count() <= 0) return nullptr;
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  static bool pool_ready = false;
  if (!pool_ready) { dev_pool_setup(); pool_ready = true; }
  MBuf *b = alloc_buf(nbytes);
  if (!b) return nullptr;
  protect(b, H0);
  b->dev_valid = 0; b->refs = 1;
  g_live[(uintptr_t)b->host] = b;
  return b->host;
}

void *pdlb200_mbuf_adopt(const void *src, size_t nbytes, char *err, size_t errlen) {
  Err E{err, errlen};
  void *h = pdlb200_mbuf_new(nbytes);
  if (!h) { E.fail(PDLB200_ECUDA, "pdlb200_mbuf_adopt: cannot allocate %zu bytes", nbytes); return nullptr; }
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  MBuf *b = find_base(h);
  if (nbytes) {
    if (dev_h2d(b->dev, src, nbytes, nullptr) != 0) { E.fail(PDLB200_ECUDA, "pdlb200_mbuf_adopt: %s", dev_err()); pdlb200_mbuf_free(h); return nullptr; }
    g_stat[2]++; g_stat[3] += nbytes;
  }
  b->dev_valid = 1;
  g_stat[7]++;
  return h;
}

void pdlb200_mbuf_retain(void *host) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (MBuf *b = find_base(host)) b->refs++;
}

void pdlb200_mbuf_free(void *host) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  MBuf *b = find_base(host);
  if (!b || --b->refs > 0) return;
  g_live.erase((uintptr_t)host);
  protect(b, H0);                                // a recycled buffer starts protected; its pages stay for the next download
  b->dev_valid = 0;
  if (g_cached + b->span <= cache_cap()) { g_free.emplace(b->nbytes, b); g_cached += b->span; return; }
  destroy_buf(b);
}

int pdlb200_mbuf_is(const void *host) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  return find_base(host) != nullptr;
}

void *pdlb200_mbuf_dev(void *host, int for_write, int discard, void *stream, char *err, size_t errlen) {
  Err E{err, errlen};
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  MBuf *b = find_base(host);
  if (!b) { E.fail(PDLB200_EINVAL, "pdlb200_mbuf_dev: %p is not a store buffer", host); return nullptr; }
  if (!b->dev_valid && !discard && b->hstate != H0 && b->nbytes) {
    if (dev_h2d(b->dev, b->alias, b->nbytes, stream) != 0) { E.fail(PDLB200_ECUDA, "pdlb200_mbuf_dev: upload: %s", dev_err()); return nullptr; }
    g_stat[2]++; g_stat[3] += b->nbytes;
  }
  b->dev_valid = 1;
  if (for_write) protect(b, H0);                 // the mirror is stale from here on
  else if (b->hstate == HW) protect(b, HR);      // both copies valid: a later host WRITE must fault to be noticed
  return b->dev;
}

int pdlb200_mbuf_host(void *host, int for_write, char *err, size_t errlen) {
  Err E{err, errlen};
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  MBuf *b = find_base(host);
  if (!b) return PDLB200_OK;                     // not ours: plain host memory is always current
  if (to_host(b, for_write != 0) != 0) return E.fail(PDLB200_ECUDA, "pdlb200_mbuf_host: download of %zu bytes failed", b->nbytes);
  return PDLB200_OK;
}

int pdlb200_mbuf_state(const void *host) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  MBuf *b = find_base(host);
  return b ? (b->hstate | (b->dev_valid ? 4 : 0)) : -1;
}

void pdlb200_mbuf_stats(uint64_t *out) {
  for (int i = 0; i < 8; i++) out[i] = g_stat[i].load();
}

// device-op registry: which transformation vtables of the host core run on the device (the binding registers
// them at attach time; its make_trans_mutual wrapper asks).  Plain pointers, no host-core types.
static std::unordered_map<const void *, int> g_devops;
void pdlb200_devop_register(const void *vtable, int on) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (on) g_devops[vtable] = 1; else g_devops.erase(vtable);
}
int pdlb200_devop_is(const void *vtable) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  return g_devops.count(vtable) != 0;
}

void pdlb200_mbuf_trim(void) {
  std::vector<MBuf *> drop;
  {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    for (auto &kv : g_free) drop.push_back(kv.second);
    g_free.clear(); g_cached = 0;
  }
  for (MBuf *d : drop) destroy_buf(d);
}

}  // extern "C"
