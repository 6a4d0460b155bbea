// scan_onepass.cuh — single-pass scan of few very long rows (included by scan.cu; cumusumover of a 1-D
// ndarray is the common case, lib/PDL/Ufunc.pd:120-141).  Ideal traffic: every element is read once and
// written once (the three-pass chunked path reads it twice).
//
// One persistent CTA per SM, 23 warps with fixed roles around a ring of OP_SLOTS shared-memory tiles of 40 KB;
// tiles move in AND out with bulk-async copies (cp.async.bulk, UBLKCP), the scan happens in place in shared memory:
//   producer (1 thread)  takes the next tile number from a global counter (tiles are handed out in the order
//                        CTAs actually run, so a look-back never waits on a CTA that is not resident) and
//                        fills the slot with ONE bulk copy armed on an mbarrier (expect_tx);
//   aggregators (4 warps) run AHEAD of the scanners: per-segment totals of the tile, the tile aggregate, and
//                        the tile's descriptor {status A, aggregate} published to global;
//   prefix warp          decoupled look-back over the 32 preceding descriptors per round trip until it meets
//                        an inclusive prefix (status P), then publishes {P, prefix+aggregate} — before the
//                        tile itself is scanned, so successors are released early;
//   scanners (16 warps)  one segment each; every LANE owns OP_VPL consecutive 16-byte vectors (5: an odd lane
//                        stride of 80 bytes keeps the 128-bit shared-memory accesses bank-conflict free): a
//                        serial in-lane scan, ONE warp shuffle scan of the lane totals per tile, results written
//                        back in place (≈ 0.1 warp instructions per element);
//   storer (1 thread)    ONE bulk copy shared -> global per tile; the slot is free again once it has been read.
// Descriptors are one 64-bit word {status, value} for 4-byte results and one 16-byte vector {status, value}
// for 8-byte results; the array and the counter are cleared by one memset per launch.
#pragma once

namespace pdlb200 {

constexpr int OP_NSCAN = 16;
constexpr int OP_NAGG = 4;
constexpr int OP_VPL = 5;                                        // 16-byte vectors per scanner lane
constexpr int OP_SEG_VECS = 32 * OP_VPL;                         // 160 vectors per segment
constexpr int OP_TILE_BYTES = OP_NSCAN * OP_SEG_VECS * 16;       // 40960
constexpr int OP_SLOTS = 5;
constexpr int OP_THREADS = (OP_NSCAN + OP_NAGG + 3) * 32;

struct OpPlan {
  const char *a; char *b;
  unsigned long long *desc;
  unsigned int *counter;
  int64_t n, tpr, ntiles, sa, sb;
  uint64_t abad, bbad;
  int abadnan, badmode;
};

__device__ __forceinline__ void op_mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void op_mbar_arrive(uint64_t *bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" :: "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void op_mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n"
               :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void op_mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nOPW_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra OPD_%=;\nbra OPW_%=;\nOPD_%=:\n}\n"
      :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void op_bulk_store(void *dst, const void *src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
               :: "l"(dst), "r"((unsigned)__cvta_generic_to_shared(src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void op_bulk_load(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes),
                  "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// status: 0 = not yet, 1 = aggregate of the tile alone, 2 = inclusive prefix of the row up to and with the tile
template <class O, int SZ = sizeof(O)> struct OpDesc;
template <class O> struct OpDesc<O, 4> {
  static constexpr int BYTES = 8;
  static __device__ __forceinline__ void put(unsigned long long *d, int64_t t, unsigned status, O v) {
    unsigned bits; memcpy(&bits, &v, 4);
    const unsigned long long w = ((unsigned long long)status << 32) | bits;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;\n" :: "l"(d + t), "l"(w) : "memory");
  }
  static __device__ __forceinline__ unsigned get(const unsigned long long *d, int64_t t, O &v) {
    unsigned long long w;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(w) : "l"(d + t) : "memory");
    const unsigned bits = (unsigned)w; memcpy(&v, &bits, 4);
    return (unsigned)(w >> 32);
  }
};
template <class O> struct OpDesc<O, 8> {
  static constexpr int BYTES = 16;
  static __device__ __forceinline__ void put(unsigned long long *d, int64_t t, unsigned status, O v) {
    unsigned long long bits; memcpy(&bits, &v, 8);
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};\n" :: "l"(d + 2 * t), "l"((unsigned long long)status), "l"(bits) : "memory");
  }
  static __device__ __forceinline__ unsigned get(const unsigned long long *d, int64_t t, O &v) {
    unsigned long long st, bits;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];\n" : "=l"(st), "=l"(bits) : "l"(d + 2 * t) : "memory");
    memcpy(&v, &bits, 8);
    return (unsigned)st;
  }
};

template <class T, class O, bool PROD, bool BAD>
__global__ void __launch_bounds__(OP_THREADS, 1) scan_onepass_kernel(const __grid_constant__ OpPlan p) {
  extern __shared__ __align__(128) unsigned char op_tiles[];
  __shared__ struct {
    O segtot[OP_SLOTS][OP_NSCAN], segpre[OP_SLOTS][OP_NSCAN];
    O tagg[OP_SLOTS], tpre[OP_SLOTS];
    int stile[OP_SLOTS];
    uint64_t full[OP_SLOTS], aggd[OP_SLOTS], pref[OP_SLOTS], scanned[OP_SLOTS], empty[OP_SLOTS];
  } c;
  constexpr int TE = OP_TILE_BYTES / (int)sizeof(T);
  constexpr int VEC = 16 / (int)sizeof(T);
  const T abad = from_bits<T>(p.abad);
  const O bbad = from_bits<O>(p.bbad);
  const O ident = PROD ? O(1) : O(0);
  const int lane = threadIdx.x & 31;
  // broadcast from lane 0: tells the compiler the role (and every tile number below) is warp-uniform, so the
  // shuffles in the role loops need no reconvergence bookkeeping
  const int wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const bool badnan = p.abadnan != 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < OP_SLOTS; s++) {
      op_mbar_init(&c.full[s], 1); op_mbar_init(&c.aggd[s], 1); op_mbar_init(&c.pref[s], 1);
      op_mbar_init(&c.scanned[s], OP_NSCAN); op_mbar_init(&c.empty[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (wid == OP_NSCAN + OP_NAGG + 1) {
    // ---- producer ----
    if (lane != 0) return;
    // the tile number is fetched one tile ahead, so the atomic's round trip is not in series with the slot wait
    long long t = (long long)atomicAdd(p.counter, 1u);
    for (unsigned it = 0;; it++) {
      const int s = it % OP_SLOTS;
      const unsigned u = it / OP_SLOTS;
      if (it >= OP_SLOTS) op_mbar_wait(&c.empty[s], (u - 1) & 1);
      if (t >= p.ntiles) { c.stile[s] = -1; op_mbar_arrive(&c.full[s]); return; }
      c.stile[s] = (int)t;
      const int64_t row = t / p.tpr, j = t - row * p.tpr;
      const int64_t left = p.n - j * TE;
      const unsigned bytes = (unsigned)((left < TE ? left : TE) * (int64_t)sizeof(T));
      op_mbar_expect_tx(&c.full[s], bytes);
      op_bulk_load(op_tiles + (size_t)s * OP_TILE_BYTES, p.a + (row * p.sa + j * TE) * (int64_t)sizeof(T), bytes, &c.full[s]);
      t = (long long)atomicAdd(p.counter, 1u);
    }
  }

  if (wid == OP_NSCAN + OP_NAGG + 2) {
    // ---- storer ----
    if (lane != 0) return;
    for (unsigned it = 0;; it++) {
      const int s = it % OP_SLOTS;
      const unsigned u = it / OP_SLOTS;
      op_mbar_wait(&c.scanned[s], u & 1);
      const long long t = c.stile[s];
      if (t < 0) break;
      const int64_t row = t / p.tpr, j = t - row * p.tpr;
      const int64_t left = p.n - j * TE;
      const unsigned bytes = (unsigned)((left < TE ? left : TE) * (int64_t)sizeof(O));
      op_bulk_store(p.b + (row * p.sb + j * TE) * (int64_t)sizeof(O), op_tiles + (size_t)s * OP_TILE_BYTES, bytes);
      asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");     // the slot has been read: reusable
      op_mbar_arrive(&c.empty[s]);
    }
    asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    return;
  }

  if (wid >= OP_NSCAN && wid < OP_NSCAN + OP_NAGG) {
    // ---- aggregators: OP_NSCAN / OP_NAGG segments per warp ----
    const int aw = wid - OP_NSCAN;
    for (unsigned it = 0;; it++) {
      const int s = it % OP_SLOTS;
      const unsigned u = it / OP_SLOTS;
      op_mbar_wait(&c.full[s], u & 1);
      const int t = __shfl_sync(0xffffffffu, c.stile[s], 0);
      if (t < 0) { if (aw == 0 && lane == 0) op_mbar_arrive(&c.aggd[s]); return; }
      const int64_t j = t % p.tpr;
      const int64_t left = p.n - j * TE;
      const int nvec = (int)((left < TE ? left : TE) / VEC);
      const uint4 *tv = reinterpret_cast<const uint4 *>(op_tiles + (size_t)s * OP_TILE_BYTES);
#pragma unroll
      for (int h = 0; h < OP_NSCAN / OP_NAGG; h++) {
        const int seg = aw * (OP_NSCAN / OP_NAGG) + h;
        O tot[OP_VPL];
#pragma unroll
        for (int k = 0; k < OP_VPL; k++) {
          tot[k] = ident;
          const int v = seg * OP_SEG_VECS + k * 32 + lane;
          if (v < nvec) {
            Pack<T> in; in.q = tv[v];
#pragma unroll
            for (int e = 0; e < VEC; e++)
              if (!(BAD && is_bad(in.e[e], abad, badnan))) tot[k] = scan_op<O, PROD>(tot[k], (O)in.e[e]);
          }
        }
        O r = tot[0];
#pragma unroll
        for (int k = 1; k < OP_VPL; k++) r = scan_op<O, PROD>(r, tot[k]);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) r = scan_op<O, PROD>(r, shfl_xor_t(r, d));
        if (lane == 0) c.segtot[s][seg] = r;
      }
      asm volatile("bar.sync 1, %0;\n" :: "n"(OP_NAGG * 32) : "memory");
      if (aw == 0 && lane == 0) {
        O agg = ident;
#pragma unroll
        for (int sg = 0; sg < OP_NSCAN; sg++) { c.segpre[s][sg] = agg; agg = scan_op<O, PROD>(agg, c.segtot[s][sg]); }
        c.tagg[s] = agg;
        OpDesc<O>::put(p.desc, t, j == 0 ? 2u : 1u, agg);
        op_mbar_arrive(&c.aggd[s]);
      }
    }
  }

  if (wid == OP_NSCAN + OP_NAGG) {
    // ---- prefix warp: decoupled look-back ----
    for (unsigned it = 0;; it++) {
      const int s = it % OP_SLOTS;
      const unsigned u = it / OP_SLOTS;
      op_mbar_wait(&c.aggd[s], u & 1);
      const int t = __shfl_sync(0xffffffffu, c.stile[s], 0);
      if (t < 0) { if (lane == 0) op_mbar_arrive(&c.pref[s]); return; }
      const int64_t j = t % p.tpr;
      O excl = ident;
      if (j > 0) {
        const int64_t lowest = t - j;             // first tile of this row: always published as P
        int64_t pos = (int64_t)t - 1;
        long long t0 = 0; bool timing = false;
        for (;;) {
          const int64_t idx = pos - lane;
          const bool valid = idx >= lowest;
          unsigned st = 1; O v = ident;
          if (valid) st = OpDesc<O>::get(p.desc, idx, v);
          const unsigned pm = __ballot_sync(0xffffffffu, valid && st == 2);
          const unsigned zm = __ballot_sync(0xffffffffu, valid && st == 0);
          const int np = pm ? __ffs(pm) - 1 : 32;
          const unsigned need = np < 31 ? ((2u << np) - 1u) : 0xffffffffu;
          if (zm & need) {
            // a predecessor has not published yet: it is resident (tiles are taken in execution order)
            if (!timing) { timing = true; t0 = clock64(); }
            else if (clock64() - t0 > 20000000000ll) __trap();     // ≈10 s: fail loudly instead of hanging the GPU
            __nanosleep(20);
            continue;
          }
          O x = (valid && lane <= np) ? v : ident;
#pragma unroll
          for (int d = 16; d >= 1; d >>= 1) x = scan_op<O, PROD>(x, shfl_xor_t(x, d));
          excl = scan_op<O, PROD>(x, excl);
          if (np < 32) break;
          pos -= 32;
        }
      }
      if (lane == 0) {
        c.tpre[s] = excl;
        if (j > 0) OpDesc<O>::put(p.desc, t, 2u, scan_op<O, PROD>(excl, c.tagg[s]));
        op_mbar_arrive(&c.pref[s]);
      }
      __syncwarp();
    }
  }

  // ---- scanners: warp `wid` owns segment `wid` of every tile, lane owns OP_VPL consecutive vectors of it ----
  for (unsigned it = 0;; it++) {
    const int s = it % OP_SLOTS;
    const unsigned u = it / OP_SLOTS;
    op_mbar_wait(&c.pref[s], u & 1);
    op_mbar_wait(&c.full[s], u & 1);                // already complete: makes the bulk copy's writes visible to this warp
    const int t = __shfl_sync(0xffffffffu, c.stile[s], 0);
    if (t < 0) { if (lane == 0) op_mbar_arrive(&c.scanned[s]); return; }
    const int64_t j = t % p.tpr;
    const int64_t left = p.n - j * TE;
    const int nvec = (int)((left < TE ? left : TE) / VEC);
    const int v0 = wid * OP_SEG_VECS + lane * OP_VPL;
    if (wid * OP_SEG_VECS < nvec) {
      const O carry = scan_op<O, PROD>(c.tpre[s], c.segpre[s][wid]);
      uint4 *tv = reinterpret_cast<uint4 *>(op_tiles + (size_t)s * OP_TILE_BYTES);
      Pack<T> in[OP_VPL];
#pragma unroll
      for (int k = 0; k < OP_VPL; k++) if (v0 + k < nvec) in[k].q = tv[v0 + k];
      O x[OP_VPL][VEC];
      unsigned bdm = 0;
      O run = ident;
#pragma unroll
      for (int k = 0; k < OP_VPL; k++) {
        const bool inr = v0 + k < nvec;
#pragma unroll
        for (int e = 0; e < VEC; e++) {
          const T val = inr ? in[k].e[e] : T(0);
          const bool bd = BAD && inr && is_bad(val, abad, badnan);
          if (BAD && bd) bdm |= 1u << (k * VEC + e);
          if (inr && !bd) run = scan_op<O, PROD>(run, (O)val);
          x[k][e] = run;
        }
      }
      O pre = run;                                    // inclusive scan of the lane totals
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const O y = shfl_up_t(pre, d);
        if (lane >= d) pre = scan_op<O, PROD>(y, pre);
      }
      O excl = shfl_up_t(pre, 1);
      if (lane == 0) excl = ident;
      const O base = scan_op<O, PROD>(carry, excl);
#pragma unroll
      for (int k = 0; k < OP_VPL; k++) {
        if (v0 + k < nvec) {
          Pack<O> out;
#pragma unroll
          for (int e = 0; e < VEC; e++)
            out.e[e] = (BAD && ((bdm >> (k * VEC + e)) & 1u)) ? bbad : scan_op<O, PROD>(base, x[k][e]);
          tv[v0 + k] = out.q;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> visible to the bulk store
    }
    __syncwarp();
    if (lane == 0) op_mbar_arrive(&c.scanned[s]);
  }
}

// Returns true when the launch was taken (rc holds the status); false: not eligible, use the three-pass path.
template <class T, class O, bool PROD>
static bool scan_onepass_try(const ScPlan &p, cudaStream_t s, const char *name, const Err &E, int *rc) {
  if constexpr (sizeof(T) != sizeof(O) || sizeof(T) < 4) return false;
  else {
    static const bool off = [] { const char *e = getenv("PDLB200_SCAN"); return e && !strcmp(e, "3pass"); }();
    if (off) return false;
    constexpr int64_t TE = OP_TILE_BYTES / (int64_t)sizeof(T);
    if (p.inc_a != 1 || p.inc_b != 1 || p.nd > 1) return false;
    if ((p.n * (int64_t)sizeof(T)) % 16 != 0 || p.n < 16 * TE) return false;
    if (((uintptr_t)p.a & 15) || ((uintptr_t)p.b & 15)) return false;
    int64_t sa = 0, sb = 0;
    if (p.nd == 1) {
      sa = p.sa[0]; sb = p.sb[0];
      if ((sa * (int64_t)sizeof(T)) % 16 != 0 || (sb * (int64_t)sizeof(O)) % 16 != 0 || sa < 0 || sb < 0) return false;
    } else if (p.nrows != 1) return false;
    const int64_t tpr = (p.n + TE - 1) / TE;
    const int64_t ntiles = tpr * p.nrows;
    if (ntiles < 2 * (int64_t)sm_count() || ntiles > (1ll << 30)) return false;
    OpPlan q;
    memset(&q, 0, sizeof q);
    q.a = p.a; q.b = p.b; q.n = p.n; q.tpr = tpr; q.ntiles = ntiles; q.sa = sa; q.sb = sb;
    q.abad = p.abad; q.bbad = p.bbad; q.abadnan = p.abadnan; q.badmode = p.badmode;
    const size_t dbytes = (size_t)ntiles * OpDesc<O>::BYTES;
    char *scr = (char *)scratch(dbytes + 16, s);
    if (!scr) { *rc = E.fail(PDLB200_ECUDA, "%s: cannot allocate scan scratch", name); return true; }
    q.desc = (unsigned long long *)scr;
    q.counter = (unsigned int *)(scr + dbytes);
    if (cudaMemsetAsync(scr, 0, dbytes + 16, s) != cudaSuccess) { *rc = E.fail(PDLB200_ECUDA, "%s: memset failed", name); return true; }
    static const bool attr = [] {
      return cudaFuncSetAttribute(scan_onepass_kernel<T, O, PROD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  OP_SLOTS * OP_TILE_BYTES) == cudaSuccess &&
             cudaFuncSetAttribute(scan_onepass_kernel<T, O, PROD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  OP_SLOTS * OP_TILE_BYTES) == cudaSuccess;
    }();
    if (!attr) { cudaGetLastError(); return false; }
    const int64_t g = ntiles < sm_count() ? ntiles : sm_count();
    if (q.badmode) scan_onepass_kernel<T, O, PROD, true><<<(int)g, OP_THREADS, OP_SLOTS * OP_TILE_BYTES, s>>>(q);
    else scan_onepass_kernel<T, O, PROD, false><<<(int)g, OP_THREADS, OP_SLOTS * OP_TILE_BYTES, s>>>(q);
    note_launch(name);
    cudaError_t e = cudaGetLastError();
    *rc = e == cudaSuccess ? PDLB200_OK : E.fail(PDLB200_ECUDA, "%s: %s", name, cudaGetErrorString(e));
    return true;
  }
}

}  // namespace pdlb200
