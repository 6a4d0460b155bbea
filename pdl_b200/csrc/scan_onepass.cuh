// scan_onepass.cuh — single-pass scan of few very long rows (included by scan.cu; cumusumover of a 1-D
// ndarray is the common case, lib/PDL/Ufunc.pd:120-141).  Ideal traffic: every element is read once and
// written once (the three-pass chunked path reads it twice).
//
// One persistent CTA per SM, 19 warps with fixed roles around a ring of 6 shared-memory tiles of 36 KB; tiles move
// in AND out with bulk-async copies (cp.async.bulk, UBLKCP), the scan happens in place in shared memory:
//   producer (1 thread)  takes the next tile number from a global counter when a slot is free (tiles are handed out
//                        in the order CTAs actually run, so a look-back never waits on a CTA that is not resident)
//                        and fills the slot with ONE bulk copy armed on an mbarrier (expect_tx);
//   aggregators (8 warps) run AHEAD of the scanners: per-segment totals of the tile, the tile aggregate, and
//                        the tile's descriptor {status A, aggregate} published to global;
//   prefix warp          decoupled look-back over the 32 preceding descriptors per round trip (the first poll is
//                        issued while the aggregators still work on the tile) until it meets an inclusive prefix
//                        (status P), then {P, prefix+aggregate} published — before the tile itself is scanned,
//                        so successors are released early;
//   scanners (8 warps)   one segment each; every LANE owns OP_VPL consecutive 16-byte vectors (9: an odd lane
//                        stride keeps the 128-bit shared-memory accesses bank-conflict free): a serial in-lane
//                        scan, ONE warp shuffle scan of the lane totals per tile, results written back in place;
//   storer (1 thread)    ONE bulk copy shared -> global per tile; the slot is free again once it has been read.
// Descriptors are one 64-bit word {status, value} for 4-byte results and one 16-byte vector {status, value}
// for 8-byte results; the array and the counter are cleared by one memset per launch.
// Tuning notes (B200, stage timers below): the ring is latency-bound — a slot is held from the
// load's issue until its prefix is known and the tile is stored — so what pays is more slots in flight and fewer
// warps contending for issue slots (16 scanner warps x 5 slots of 40 KB: 0.82 of the copy peak; 8 x 6 x 36 KB: 0.88);
// wider look-back windows (2..8 x 32 descriptors per poll) and several prefix warps per CTA both made it slower.
#pragma once

namespace pdlb200 {

#ifndef OP_NSCAN_V
#define OP_NSCAN_V 8
#endif
#ifndef OP_VPL_V
#define OP_VPL_V 9
#endif
#ifndef OP_SLOTS_V
#define OP_SLOTS_V 6
#endif
constexpr int OP_NSCAN = OP_NSCAN_V;
#ifndef OP_NAGG_V
#define OP_NAGG_V 8
#endif
constexpr int OP_NAGG = OP_NAGG_V;
constexpr int OP_VPL = OP_VPL_V;                                        // 16-byte vectors per scanner lane
constexpr int OP_SEG_VECS = 32 * OP_VPL;                         // 96 vectors per segment
constexpr int OP_TILE_BYTES = OP_NSCAN * OP_SEG_VECS * 16;       // 24576
constexpr int OP_SLOTS = OP_SLOTS_V;
constexpr int OP_NPREF = 1;                                      // prefix warps: look-backs in flight per SM
constexpr int OP_THREADS = (OP_NSCAN + OP_NAGG + OP_NPREF + 2) * 32;

// Stage timers for tuning (clock64 sums per CTA, 16 slots): compiled in only with -DPDLB200_SCAN_PROF
#ifdef PDLB200_SCAN_PROF
#define OP_PROF_DECL long long prof_acc[16] = {0}; (void)prof_acc;
#define OP_PROF_T(v) const long long v = clock64();
#define OP_PROF_ADD(k, v) prof_acc[k] += clock64() - v;
#define OP_PROF_INC(k) prof_acc[k] += 1;
#define OP_PROF_OUT(k) if (p.prof) p.prof[blockIdx.x * 16 + k] = prof_acc[k];
#define OP_PROF_STAMP(s) c.stamp[s] = clock64();
#define OP_PROF_SINCE(k, s) prof_acc[k] += clock64() - c.stamp[s];
#else
#define OP_PROF_DECL
#define OP_PROF_T(v)
#define OP_PROF_ADD(k, v)
#define OP_PROF_INC(k)
#define OP_PROF_OUT(k)
#define OP_PROF_STAMP(s)
#define OP_PROF_SINCE(k, s)
#endif

struct OpPlan {
  long long *prof;
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

// BADK: 0 = no BAD values, 1 = BAD by value, 2 = BAD is NaN (compile-time: one compare per element in the hot loops)
template <class T, class O, bool PROD, int BADK>
__global__ void __launch_bounds__(OP_THREADS, 1) scan_onepass_kernel(const __grid_constant__ OpPlan p) {
  extern __shared__ __align__(128) unsigned char op_tiles[];
  __shared__ struct {
    O segtot[OP_SLOTS][OP_NSCAN], segpre[OP_SLOTS][OP_NSCAN];
    O tagg[OP_SLOTS], tpre[OP_SLOTS];
    int stile[OP_SLOTS];
#ifdef PDLB200_SCAN_PROF
    long long stamp[OP_SLOTS];
#endif
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
  constexpr bool BAD = BADK != 0, badnan = BADK == 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < OP_SLOTS; s++) {
      op_mbar_init(&c.full[s], 1); op_mbar_init(&c.aggd[s], 1); op_mbar_init(&c.pref[s], 1);
      op_mbar_init(&c.scanned[s], OP_NSCAN); op_mbar_init(&c.empty[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (wid == OP_NSCAN + OP_NAGG + OP_NPREF) {
    // ---- producer ----
    if (lane != 0) return;
    OP_PROF_DECL
    for (unsigned it = 0;; it++) {
      const int s = it % OP_SLOTS;
      const unsigned u = it / OP_SLOTS;
      OP_PROF_T(a0)
      if (it >= OP_SLOTS) op_mbar_wait(&c.empty[s], (u - 1) & 1);
      OP_PROF_ADD(0, a0)
      // taken only when the slot is free: a tile claimed early but loaded late would stall every successor's look-back
      OP_PROF_T(a1)
      const long long t = (long long)atomicAdd(p.counter, 1u);
      OP_PROF_ADD(1, a1)
      if (t >= p.ntiles) {
        // end of work, posted for OP_NPREF consecutive iterations: every prefix warp meets it in its own sequence
        c.stile[s] = -1; op_mbar_arrive(&c.full[s]);
        for (unsigned e = 1; e < OP_NPREF; e++) {
          const unsigned it2 = it + e;
          const int s2 = it2 % OP_SLOTS;
          if (it2 >= OP_SLOTS) op_mbar_wait(&c.empty[s2], (it2 / OP_SLOTS - 1) & 1);
          c.stile[s2] = -1; op_mbar_arrive(&c.full[s2]);
        }
        OP_PROF_OUT(0) OP_PROF_OUT(1)
        return;
      }
      c.stile[s] = (int)t;
      OP_PROF_STAMP(s)
      const int64_t row = t / p.tpr, j = t - row * p.tpr;
      const int64_t left = p.n - j * TE;
      const unsigned bytes = (unsigned)((left < TE ? left : TE) * (int64_t)sizeof(T));
      op_mbar_expect_tx(&c.full[s], bytes);
      op_bulk_load(op_tiles + (size_t)s * OP_TILE_BYTES, p.a + (row * p.sa + j * TE) * (int64_t)sizeof(T), bytes, &c.full[s]);
    }
  }

  if (wid == OP_NSCAN + OP_NAGG + OP_NPREF + 1) {
    // ---- storer ----
    if (lane != 0) return;
    OP_PROF_DECL
    for (unsigned it = 0;; it++) {
      const int s = it % OP_SLOTS;
      const unsigned u = it / OP_SLOTS;
      OP_PROF_T(a9)
      op_mbar_wait(&c.scanned[s], u & 1);
      OP_PROF_ADD(9, a9)
      const long long t = c.stile[s];
      if (t < 0) break;
      OP_PROF_T(a10)
      const int64_t row = t / p.tpr, j = t - row * p.tpr;
      const int64_t left = p.n - j * TE;
      const unsigned bytes = (unsigned)((left < TE ? left : TE) * (int64_t)sizeof(O));
      op_bulk_store(p.b + (row * p.sb + j * TE) * (int64_t)sizeof(O), op_tiles + (size_t)s * OP_TILE_BYTES, bytes);
      asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");     // the slot has been read: reusable
      OP_PROF_ADD(10, a10)
      op_mbar_arrive(&c.empty[s]);
    }
    asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    OP_PROF_OUT(9) OP_PROF_OUT(10)
    return;
  }

  if (wid >= OP_NSCAN && wid < OP_NSCAN + OP_NAGG) {
    // ---- aggregators: OP_NSCAN / OP_NAGG segments per warp ----
    const int aw = wid - OP_NSCAN;
    int nend = 0;
    OP_PROF_DECL
    for (unsigned it = 0;; it++) {
      const int s = it % OP_SLOTS;
      const unsigned u = it / OP_SLOTS;
      OP_PROF_T(a11)
      op_mbar_wait(&c.full[s], u & 1);
      OP_PROF_ADD(11, a11)
      const int t = __shfl_sync(0xffffffffu, c.stile[s], 0);
      if (t < 0) {
        if (aw == 0 && lane == 0) op_mbar_arrive(&c.aggd[s]);
        if (++nend < OP_NPREF) continue;
        if (aw == 0 && lane == 0) { OP_PROF_OUT(11) OP_PROF_OUT(2) OP_PROF_OUT(3) }
        return;
      }
      OP_PROF_SINCE(2, s)
      OP_PROF_T(a3)
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
        OP_PROF_ADD(3, a3)
      }
    }
  }

  if (wid >= OP_NSCAN + OP_NAGG && wid < OP_NSCAN + OP_NAGG + OP_NPREF) {
    // ---- prefix warps: decoupled look-back, warp k takes every OP_NPREF-th tile of this CTA ----
    OP_PROF_DECL
    for (unsigned it = wid - (OP_NSCAN + OP_NAGG);; it += OP_NPREF) {
      const int s = it % OP_SLOTS;
      const unsigned u = it / OP_SLOTS;
      // the tile number is known once the tile has landed: poll the 32 predecessors NOW, while the aggregators
      // are still working on the tile — one L2 round trip (≈ 1000 cycles under load) off the serial path
      op_mbar_wait(&c.full[s], u & 1);
      const int t = __shfl_sync(0xffffffffu, c.stile[s], 0);
      int64_t j = 0, lowest = 0, pos = 0;
      unsigned st = 1; O v = ident;
      if (t >= 0) {
        j = t % p.tpr; lowest = t - j; pos = (int64_t)t - 1;       // tile `lowest` starts the row: always published as P
        if (j > 0 && pos - lane >= lowest) st = OpDesc<O>::get(p.desc, pos - lane, v);
      }
      OP_PROF_T(a12)
      op_mbar_wait(&c.aggd[s], u & 1);
      OP_PROF_ADD(12, a12)
      if (t < 0) { if (lane == 0) { op_mbar_arrive(&c.pref[s]); if (wid == OP_NSCAN + OP_NAGG) { OP_PROF_OUT(12) OP_PROF_OUT(4) OP_PROF_OUT(5) OP_PROF_OUT(6) } } return; }
      OP_PROF_T(a4)
      O excl = ident;
      if (j > 0) {
        long long t0 = 0; bool timing = false;
        for (bool first = true;; first = false) {
          if (!first) { st = 1; v = ident; if (pos - lane >= lowest) st = OpDesc<O>::get(p.desc, pos - lane, v); }
          const unsigned pm = __ballot_sync(0xffffffffu, st == 2);
          const unsigned zm = __ballot_sync(0xffffffffu, st == 0);
          const int np = pm ? __ffs(pm) - 1 : 32;
          const unsigned need = np < 31 ? ((2u << np) - 1u) : 0xffffffffu;
          OP_PROF_INC(5)
          if (zm & need) {
            OP_PROF_INC(6)
            // a predecessor has not published yet: it is resident (tiles are taken in execution order)
            if (!timing) { timing = true; t0 = clock64(); }
            else if (clock64() - t0 > 20000000000ll) __trap();     // ≈10 s: fail loudly instead of hanging the GPU
            __nanosleep(20);
            continue;
          }
          O x = lane <= np ? v : ident;
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
      OP_PROF_ADD(4, a4)
      __syncwarp();
    }
  }

  // ---- scanners: warp `wid` owns segment `wid` of every tile, lane owns OP_VPL consecutive vectors of it ----
  OP_PROF_DECL
  for (unsigned it = 0;; it++) {
    const int s = it % OP_SLOTS;
    const unsigned u = it / OP_SLOTS;
    OP_PROF_T(a7)
    op_mbar_wait(&c.pref[s], u & 1);
    op_mbar_wait(&c.full[s], u & 1);                // already complete: makes the bulk copy's writes visible to this warp
    OP_PROF_ADD(7, a7)
    const int t = __shfl_sync(0xffffffffu, c.stile[s], 0);
    if (t < 0) { if (lane == 0) { op_mbar_arrive(&c.scanned[s]); if (wid == 0) { OP_PROF_OUT(7) OP_PROF_OUT(8) } } return; }
    OP_PROF_T(a8)
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
      using Mask = typename std::conditional<(OP_VPL * VEC > 32), unsigned long long, unsigned>::type;
      Mask bdm = 0;
      O run = ident;
#pragma unroll
      for (int k = 0; k < OP_VPL; k++) {
        const bool inr = v0 + k < nvec;
#pragma unroll
        for (int e = 0; e < VEC; e++) {
          const T val = inr ? in[k].e[e] : T(0);
          const bool bd = BAD && inr && is_bad(val, abad, badnan);
          if (BAD && bd) bdm |= Mask(1) << (k * VEC + e);
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
            out.e[e] = (BAD && ((bdm >> (k * VEC + e)) & Mask(1))) ? bbad : scan_op<O, PROD>(base, x[k][e]);
          tv[v0 + k] = out.q;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> visible to the bulk store
    }
    __syncwarp();
    OP_PROF_ADD(8, a8)
    if (lane == 0) op_mbar_arrive(&c.scanned[s]);
  }
}

// rows whose length is not a whole number of 16-byte vectors: the (< 16 bytes of) elements after the last vector,
// one thread per row, carried on from the row's inclusive prefix — the P descriptor of the row's last tile
template <class T, class O, bool PROD>
__global__ void scan_onepass_tail_kernel(const __grid_constant__ OpPlan p, int64_t n_all, int64_t nrows) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  const T abad = from_bits<T>(p.abad);
  const O bbad = from_bits<O>(p.bbad);
  O carry;
  OpDesc<O>::get(p.desc, row * p.tpr + p.tpr - 1, carry);
  const T *a = reinterpret_cast<const T *>(p.a) + row * p.sa;
  O *b = reinterpret_cast<O *>(p.b) + row * p.sb;
  for (int64_t k = p.n; k < n_all; k++) {
    const T v = a[k];
    if (p.badmode && is_bad(v, abad, p.abadnan != 0)) { b[k] = bbad; continue; }
    carry = scan_op<O, PROD>(carry, (O)v);
    b[k] = carry;
  }
}

// Returns true when the launch was taken (rc holds the status); false: not eligible, use the three-pass path.
template <class T, class O, bool PROD>
static bool scan_onepass_try(const ScPlan &p, cudaStream_t s, const char *name, const Err &E, int *rc) {
  if constexpr (sizeof(T) != sizeof(O) || sizeof(T) < 4) return false;
  else {
    { const char *e = getenv("PDLB200_SCAN"); if (e && !strcmp(e, "3pass")) return false; }
    constexpr int64_t TE = OP_TILE_BYTES / (int64_t)sizeof(T);
    constexpr int64_t VEC = 16 / (int64_t)sizeof(T);
    if (p.inc_a != 1 || p.inc_b != 1 || p.nd > 1) return false;
    const int64_t nmain = p.n - p.n % VEC;          // whole 16-byte vectors: the bulk copies' granularity; the rest is the tail
    if (nmain < 16 * TE) return false;
    if (((uintptr_t)p.a & 15) || ((uintptr_t)p.b & 15)) return false;
    int64_t sa = 0, sb = 0;
    if (p.nd == 1) {
      sa = p.sa[0]; sb = p.sb[0];
      if ((sa * (int64_t)sizeof(T)) % 16 != 0 || (sb * (int64_t)sizeof(O)) % 16 != 0 || sa < 0 || sb < 0) return false;
    } else if (p.nrows != 1) return false;
    const int64_t tpr = (nmain + TE - 1) / TE;
    const int64_t ntiles = tpr * p.nrows;
    if (ntiles < 2 * (int64_t)sm_count() || ntiles > (1ll << 30)) return false;
    OpPlan q;
    memset(&q, 0, sizeof q);
    q.a = p.a; q.b = p.b; q.n = nmain; q.tpr = tpr; q.ntiles = ntiles; q.sa = sa; q.sb = sb;
    q.abad = p.abad; q.bbad = p.bbad; q.abadnan = p.abadnan; q.badmode = p.badmode;
    const size_t dbytes = (size_t)ntiles * OpDesc<O>::BYTES;
    char *scr = (char *)scratch(dbytes + 16, s);
    if (!scr) { *rc = E.fail(PDLB200_ECUDA, "%s: cannot allocate scan scratch", name); return true; }
    q.desc = (unsigned long long *)scr;
    q.counter = (unsigned int *)(scr + dbytes);
    if (cudaMemsetAsync(scr, 0, dbytes + 16, s) != cudaSuccess) { *rc = E.fail(PDLB200_ECUDA, "%s: memset failed", name); return true; }
    static const bool attr = [] {
      bool ok = cudaFuncSetAttribute(scan_onepass_kernel<T, O, PROD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     OP_SLOTS * OP_TILE_BYTES) == cudaSuccess &&
                cudaFuncSetAttribute(scan_onepass_kernel<T, O, PROD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     OP_SLOTS * OP_TILE_BYTES) == cudaSuccess;
      if constexpr (!tt<T>::is_int)
        ok = ok && cudaFuncSetAttribute(scan_onepass_kernel<T, O, PROD, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        OP_SLOTS * OP_TILE_BYTES) == cudaSuccess;
      return ok;
    }();
    if (!attr) { cudaGetLastError(); return false; }
    const int64_t g = ntiles < sm_count() ? ntiles : sm_count();
#ifdef PDLB200_SCAN_PROF
    static long long *prof_dev = nullptr;
    if (!prof_dev) cudaMalloc(&prof_dev, 16 * 8 * 1024);
    cudaMemsetAsync(prof_dev, 0, 16 * 8 * 1024, s);
    q.prof = prof_dev;
#endif
    bool launched = false;
    if constexpr (!tt<T>::is_int) {
      if (q.badmode && q.abadnan) { scan_onepass_kernel<T, O, PROD, 2><<<(int)g, OP_THREADS, OP_SLOTS * OP_TILE_BYTES, s>>>(q); launched = true; }
    }
    if (!launched) {
      if (q.badmode) scan_onepass_kernel<T, O, PROD, 1><<<(int)g, OP_THREADS, OP_SLOTS * OP_TILE_BYTES, s>>>(q);
      else scan_onepass_kernel<T, O, PROD, 0><<<(int)g, OP_THREADS, OP_SLOTS * OP_TILE_BYTES, s>>>(q);
    }
    note_launch(name);
    if (nmain < p.n) {
      scan_onepass_tail_kernel<T, O, PROD><<<(unsigned)((p.nrows + 127) / 128), 128, 0, s>>>(q, p.n, p.nrows);
      note_launch(name);
    }
#ifdef PDLB200_SCAN_PROF
    if (getenv("PDLB200_SCAN_PROF_PRINT")) {
      static long long host[16 * 1024];
      cudaStreamSynchronize(s);
      cudaMemcpy(host, prof_dev, sizeof host, cudaMemcpyDeviceToHost);
      double sum[16] = {0};
      for (int64_t b = 0; b < g; b++) for (int k = 0; k < 16; k++) sum[k] += (double)host[b * 16 + k];
      const double tiles_per_cta = (double)ntiles / (double)g;
      fprintf(stderr, "scan_prof tiles/cta %.1f | per tile (cycles): empty_wait %.0f atomic %.0f load_lat %.0f agg %.0f "
              "lookback %.0f (rounds %.2f stalled %.2f) scan_wait %.0f scan %.0f store_wait %.0f store %.0f agg_wait_full %.0f prefix_wait_agg %.0f\n",
              tiles_per_cta, sum[0] / g / tiles_per_cta, sum[1] / g / tiles_per_cta, sum[2] / g / tiles_per_cta, sum[3] / g / tiles_per_cta,
              sum[4] * OP_NPREF / g / tiles_per_cta, sum[5] * OP_NPREF / g / tiles_per_cta, sum[6] * OP_NPREF / g / tiles_per_cta, sum[7] / g / tiles_per_cta,
              sum[8] / g / tiles_per_cta, sum[9] / g / tiles_per_cta, sum[10] / g / tiles_per_cta, sum[11] / g / tiles_per_cta, sum[12] * OP_NPREF / g / tiles_per_cta);
    }
#endif
    cudaError_t e = cudaGetLastError();
    *rc = e == cudaSuccess ? PDLB200_OK : E.fail(PDLB200_ECUDA, "%s: %s", name, cudaGetErrorString(e));
    return true;
  }
}

}  // namespace pdlb200
