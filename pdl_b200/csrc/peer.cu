// peer.cu — record exchange over peer memory (NVLink / NVSwitch P2P) for the sharded whole-array reductions.
//
// Step 2 of a sharded reduction (include/pdlb200.h PART_ / COLL_ ops) moves 32 bytes per row per rank.  Instead of a
// library collective, every rank owns a MAILBOX in its own HBM that all peers have mapped (cudaIpc handles
// exchanged once at set-up); ONE small kernel per exchange stores this rank's records straight into slot [rank] of
// every peer's mailbox over NVLink, publishes an epoch flag with release semantics at system scope, and waits
// until the flags of all peers have arrived in its own mailbox.  The COLL_* kernel that follows in stream order
// reads the gathered records from local memory.  Mailboxes are double-buffered by epoch parity: a peer can run at
// most one exchange ahead (it needs this rank's next flag to complete the one after), so the buffer the local
// COLL_* kernel is still reading is never the one being written.
//
// Mailbox layout (int64 words), fixed by the mailbox CAPACITY `cap` (words per rank), not by the size of one
// exchange — record words of a bigger earlier exchange must never sit where a later one looks for flags:
//   buffer b in {0, 1} at b * (world * cap + world):  [world][cap] records, then [world] epoch flags.
#include <cstring>
#include "common.cuh"
namespace pdlb200 {

struct PeerPlan {
  const int64_t *local;          // this rank's records, nwords int64
  int64_t *const *mailboxes;     // device array [world]: every rank's mailbox as mapped in THIS process
  int64_t nwords, cap;
  int64_t epoch;
  int rank, world;
  int *timeout_flag;
};

__global__ void __launch_bounds__(256) peer_exchange_kernel(const __grid_constant__ PeerPlan p) {
  const int64_t stride = (int64_t)p.world * p.cap + p.world;
  const int64_t boff = (p.epoch & 1) * stride;
  // send: one warp per peer (grid-stride over peers), lanes over the words
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int peer = warp; peer < p.world; peer += nwarps) {
    int64_t *dst = p.mailboxes[peer] + boff + (int64_t)p.rank * p.cap;
    for (int64_t i = lane; i < p.nwords; i += 32) dst[i] = p.local[i];
    __threadfence_system();                       // the records are visible system-wide before the flag
    __syncwarp();
    if (lane == 0) {
      int64_t *flag = p.mailboxes[peer] + boff + (int64_t)p.world * p.cap + p.rank;
      asm volatile("st.release.sys.global.u64 [%0], %1;\n" :: "l"(flag), "l"(p.epoch) : "memory");
    }
  }
  // wait: every peer's flag for this epoch in MY mailbox
  const int64_t *myflags = p.mailboxes[p.rank] + boff + (int64_t)p.world * p.cap;
  const long long t0 = clock64();
  for (int peer = threadIdx.x; peer < p.world; peer += blockDim.x) {
    int64_t v;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(myflags + peer) : "memory");
      if (v != p.epoch && clock64() - t0 > 20000000000ll) { *p.timeout_flag = 1; break; }   // ~10 s: a peer died
    } while (v != p.epoch);
  }
  __threadfence_system();
}

}  // namespace pdlb200

using namespace pdlb200;

extern "C" {

// Mailbox for `world` ranks exchanging up to `nwords` int64 words each (two epoch buffers), zero-initialised.
// Plain cudaMalloc: the allocation must be exportable with cudaIpcGetMemHandle.
void *pdlb200_peer_mailbox_new(int world, size_t cap_words) {
  if (pdlb200_device_count() <= 0 || world <= 0) return nullptr;
  const size_t bytes = 2 * ((size_t)world * cap_words + world) * sizeof(int64_t);
  void *p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  cudaMemset(p, 0, bytes);
  cudaDeviceSynchronize();
  return p;
}
void pdlb200_peer_mailbox_free(void *p) { if (p) cudaFree(p); }

int pdlb200_ipc_export(void *devptr, unsigned char *handle64, char *err, size_t errlen) {
  Err E{err, errlen};
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t h;
  PDLB200_CUDA_OK(cudaIpcGetMemHandle(&h, devptr), E);
  memcpy(handle64, &h, 64);
  return PDLB200_OK;
}
void *pdlb200_ipc_open(const unsigned char *handle64, char *err, size_t errlen) {
  Err E{err, errlen};
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void *p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { cudaGetLastError(); E.fail(PDLB200_ECUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); return nullptr; }
  return p;
}
void pdlb200_ipc_close(void *p) { if (p) cudaIpcCloseMemHandle(p); }

// One exchange: `local` (nwords int64, device) into slot [rank] of every mailbox in `mailboxes` (DEVICE array of
// `world` device pointers, this rank's own included), then wait for all peers.  `epoch` must be the same on all
// ranks and increase by 1 per exchange, starting at 1.  On return (in stream order) the records of all ranks are at
// pdlb200_peer_gathered(mailboxes[rank], ...).
int pdlb200_peer_exchange(const void *local, size_t nwords, size_t cap_words, void *const *mailboxes, int rank, int world,
                          int64_t epoch, void *stream, char *err, size_t errlen) {
  Err E{err, errlen};
  if (pdlb200_device_count() <= 0) return E.fail(PDLB200_ENODEVICE, "pdlb200_peer_exchange: no CUDA device available");
  if (!local || !mailboxes || world <= 0 || rank < 0 || rank >= world || epoch < 1 || nwords > cap_words)
    return E.fail(PDLB200_EINVAL, "pdlb200_peer_exchange: bad arguments");
  PeerPlan p;
  p.local = (const int64_t *)local; p.mailboxes = (int64_t *const *)mailboxes; p.nwords = (int64_t)nwords; p.cap = (int64_t)cap_words;
  p.epoch = epoch; p.rank = rank; p.world = world;
  cudaStream_t s = (cudaStream_t)stream;
  p.timeout_flag = (int *)scratch(64, s);
  if (!p.timeout_flag) return E.fail(PDLB200_ECUDA, "pdlb200_peer_exchange: no scratch");
  peer_exchange_kernel<<<1, 256, 0, s>>>(p);
  note_launch("peer_exchange");
  PDLB200_CUDA_OK(cudaGetLastError(), E);
  return PDLB200_OK;
}

// Word offset (int64) of the gathered [world][cap] records of `epoch` inside a mailbox.
int64_t pdlb200_peer_gathered_offset(int world, size_t cap_words, int64_t epoch) {
  return (epoch & 1) * ((int64_t)world * (int64_t)cap_words + world);
}

}  // extern "C"
