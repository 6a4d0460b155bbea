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
#include <sys/mman.h>
#include <ucontext.h>
#include <unistd.h>
#include "common.cuh"

namespace pdlb200 {
namespace {

enum HostState { H0 = 0, HR = 1, HW = 2 };

struct MBuf {
  char *host; void *dev; size_t nbytes, span;   // span = nbytes rounded up to pages (the mmap length)
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
    mprotect(b->host, b->span, PROT_READ | PROT_WRITE);
    if (b->dev_valid && b->nbytes) {
      int cur = 0; cudaGetDevice(&cur);
      if (cur != b->device) cudaSetDevice(b->device);
      // legacy default stream: ordered after every kernel the binding launched on it
      cudaError_t e = cudaMemcpy(b->host, b->dev, b->nbytes, cudaMemcpyDeviceToHost);
      if (cur != b->device) cudaSetDevice(cur);
      if (e != cudaSuccess) { mprotect(b->host, b->span, PROT_NONE); return -1; }
      g_stat[4]++; g_stat[5] += b->nbytes;
    }
    b->hstate = HW;                             // pages are RW right now
    if (!for_write) protect(b, HR);
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
  cudaGetDevice(&b->device);
  void *h = mmap(nullptr, b->span, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (h == MAP_FAILED) { delete b; return nullptr; }
  b->host = (char *)h;
  madvise(h, b->span, MADV_HUGEPAGE);            // a later download populates the mirror with 2 MiB pages where THP allows
  cudaError_t e = cudaMallocAsync(&b->dev, b->span, (cudaStream_t)0);
  if (e != cudaSuccess) {
    cudaGetLastError();
    // hand the cached pairs back and retry once
    std::vector<MBuf *> drop;
    for (auto &kv : g_free) drop.push_back(kv.second);
    g_free.clear(); g_cached = 0;
    for (MBuf *d : drop) { cudaFreeAsync(d->dev, (cudaStream_t)0); munmap(d->host, d->span); delete d; }
    cudaStreamSynchronize((cudaStream_t)0);
    e = cudaMallocAsync(&b->dev, b->span, (cudaStream_t)0);
    if (e != cudaSuccess) { cudaGetLastError(); munmap(h, b->span); delete b; return nullptr; }
  }
  g_stat[0]++;
  return b;
}

}  // namespace
}  // namespace pdlb200

using namespace pdlb200;

extern "C" {

void *pdlb200_mbuf_new(size_t nbytes) {
  if (pdlb200_device_count() <= 0) return nullptr;
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  static bool pool_ready = false;
  if (!pool_ready) {
    int dev = 0; cudaGetDevice(&dev);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) { uint64_t thr = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr); }
    pool_ready = true;
  }
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
    // pageable source: the call returns once the source has been consumed, so the caller may release it
    cudaError_t e = cudaMemcpyAsync(b->dev, src, nbytes, cudaMemcpyHostToDevice, (cudaStream_t)0);
    if (e != cudaSuccess) { E.fail(PDLB200_ECUDA, "pdlb200_mbuf_adopt: %s", cudaGetErrorString(e)); pdlb200_mbuf_free(h); return nullptr; }
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
  if (b->hstate != H0) {
    // the mirror was populated by host access: drop the pages so that a recycled buffer starts unpopulated
    madvise(b->host, b->span, MADV_DONTNEED);
    protect(b, H0);
  }
  b->dev_valid = 0;
  if (g_cached + b->span <= cache_cap()) { g_free.emplace(b->nbytes, b); g_cached += b->span; return; }
  cudaFreeAsync(b->dev, (cudaStream_t)0);
  munmap(b->host, b->span);
  delete b;
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
    cudaError_t e = cudaMemcpyAsync(b->dev, b->host, b->nbytes, cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) { E.fail(PDLB200_ECUDA, "pdlb200_mbuf_dev: upload: %s", cudaGetErrorString(e)); return nullptr; }
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

void pdlb200_mbuf_trim(void) {
  std::vector<MBuf *> drop;
  {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    for (auto &kv : g_free) drop.push_back(kv.second);
    g_free.clear(); g_cached = 0;
  }
  for (MBuf *d : drop) { cudaFreeAsync(d->dev, (cudaStream_t)0); munmap(d->host, d->span); delete d; }
}

}  // extern "C"
