// reduce.cuh — PDL::Ufunc `a(n); [o]b()` reductions on the device.
//
// Replaces the generated loop(n) inside the two broadcast `for`s of
// pdl_<op>_readdata (lib/PDL/Ufunc.pd:88-118 sumover/prodover, :143-187 and/or/..over,
// :413-444 average, :446-500 minimum/maximum/_ind).  Roofline: HBM; algorithmic
// bytes per row = n*sizeof(T) read + sizeof(O) written.
//
// A "row" is one broadcast position; its n elements sit inc_n apart.  Three
// cooperation widths, picked by the host planner (reduce_plan.cu):
//   MODE 2  one CTA per row     (long rows: 128-bit loads, UNROLL in flight, smem tree)
//   MODE 1  one warp per row    (medium rows: shuffle tree only, no barrier)
//   MODE 0  one thread per row  (short rows, and "column" reductions where a broadcast
//                                dim is the unit-stride one: adjacent threads read
//                                adjacent rows, so the loads coalesce across the warp)
// When there are too few rows to fill 148 SMs each row is cut into chunks
// (blockIdx.y); partial accumulators go to scratch and a second tiny kernel merges them.
//
// Every reducer has two accumulators: a lean per-thread `Loc` that is exactly the
// reference's sequential loop body applied to that thread's elements in increasing index
// order (a few instructions per element: the kernels stay HBM-bound), and a cross-thread
// `Acc` whose merge is an order-independent restatement of the same loop (see each reducer).
// Integer, min/max and index results are therefore bit-exact however a row is cut; float
// sums differ from the reference only by summation order.
#pragma once
#include <type_traits>
#include "common.cuh"

namespace pdlb200 {

constexpr int RD_THREADS = 256;
// 128-bit loads a thread keeps in flight per trip (R::kUnroll).  HBM-bound streaming needs ~100+ KB in
// flight per SM: reducers whose hot loop fits 32 registers run 8 CTAs/SM with 4 loads each; heavier
// ones (64-bit accumulators, min/max with index) get 6 CTAs/SM and compensate with 8 loads.
constexpr int RD_UNROLL_MAX = 8;

struct RdPlan {
  const char *a; char *b;       // bases with offs applied
  int64_t n, inc_n;             // reduced dim: size, stride (elements)
  int64_t nrows;
  int64_t dims[MAXD];
  int64_t sa[MAXD], sb[MAXD];   // broadcast strides of a and b (elements)
  int64_t chunk;                // elements of n per chunk (== n when nchunks == 1); always < 2^31
  int64_t goff, inc_r;          // PART_* ops: global index of element 0, stride of the record's r dim
  char *partial;                // scratch for nchunks > 1
  uint64_t abad, bbad;
  int nd;
  int nchunks;
  int abadnan;
  int badmode;                  // trans->bvalflag
};

// ---- accumulator shuffles ----------------------------------------------------
template <class A> __device__ __forceinline__ A shfl_xor_acc(const A &x, int mask) {
  static_assert(sizeof(A) % 4 == 0, "accumulator must be a multiple of 4 bytes");
  A r;
  const uint32_t *s = reinterpret_cast<const uint32_t *>(&x);
  uint32_t *d = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(A) / 4); i++) d[i] = __shfl_xor_sync(0xffffffffu, s[i], mask);
  return r;
}

template <class O> __device__ __forceinline__ O wrap_add(O a, O b) {
  if constexpr (tt<O>::is_int) { using U = typename tt<O>::wide_u; return (O)((U)a + (U)b); } else return a + b;
}
template <class O> __device__ __forceinline__ O wrap_mul(O a, O b) {
  if constexpr (tt<O>::is_int) { using U = typename tt<O>::wide_u; return (O)((U)a * (U)b); } else return a * b;
}
template <class T> __device__ __forceinline__ T nan_of() {
  if constexpr (sizeof(T) == 4) return __uint_as_float(0x7fc00000u); else return __longlong_as_double(0x7ff8000000000000ll);
}

constexpr int64_t RD_NOIDX = 0x7fffffffffffffffll;

// ---- SIMD-within-a-register helpers for 8/16-bit integer rows ---------------------------------------
// At 1-2 bytes per element a per-element loop body is issue-bound long before HBM is (5.5 issue slots per
// byte element at 6.5 TB/s), so sums of small integers work on 32-bit words: BAD lanes are found with the
// exact zero-lane test on (w ^ badword), cleared, and the word is summed by one dp4a / dp2a.
template <class T> __device__ __forceinline__ uint32_t swar_splat(T v) {
  if constexpr (sizeof(T) == 1) return 0x01010101u * (uint32_t)(uint8_t)v; else return 0x00010001u * (uint32_t)(uint16_t)v;
}
// all-ones in every lane of w that equals the lane of badw; nbad += number of such lanes
template <class T> __device__ __forceinline__ uint32_t swar_eq_mask(uint32_t w, uint32_t badw, int32_t &nbad) {
  const uint32_t x = w ^ badw;
  if constexpr (sizeof(T) == 1) {
    const uint32_t hi = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;   // 0x80 in each zero byte, exact
    nbad += __popc(hi);
    return (hi >> 7) * 0xffu;
  } else {
    const uint32_t hi = ~(((x & 0x7fff7fffu) + 0x7fff7fffu) | x) & 0x80008000u;
    nbad += __popc(hi);
    return (hi >> 15) * 0xffffu;
  }
}
// number of lanes of w equal to the lane of badw (the zero-lane test of swar_eq_mask without building the mask)
template <class T> __device__ __forceinline__ int32_t swar_eq_count(uint32_t w, uint32_t badw) {
  const uint32_t x = w ^ badw;
  if constexpr (sizeof(T) == 1) return __popc(~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u);
  else return __popc(~(((x & 0x7fff7fffu) + 0x7fff7fffu) | x) & 0x80008000u);
}
// acc + sum of the lanes of w (lanes read as T), wrapping in 32 bits
template <class T> __device__ __forceinline__ int32_t swar_sum(uint32_t w, int32_t acc) {
  if constexpr (sizeof(T) == 1) {
    if constexpr (tt<T>::is_uns) return (int32_t)__dp4a(w, 0x01010101u, (uint32_t)acc); else return __dp4a((int)w, 0x01010101, acc);
  } else {
    if constexpr (tt<T>::is_uns) return (int32_t)__dp2a_lo(w, 0x00000101u, (uint32_t)acc); else return __dp2a_lo((int)w, 0x00000101, acc);
  }
}
// reducers opt in with `static constexpr bool kPack` + `lpush_pack<BADK>(Loc &, const Pack<T> &, T abad)`
template <class R, class = void> struct rd_kpack { static constexpr bool value = false; };
template <class R> struct rd_kpack<R, std::void_t<decltype(R::kPack)>> { static constexpr bool value = R::kPack; };

// ---- reducers -------------------------------------------------------------------
// Loc: linit(), lpush(loc, value, rel) with rel = element index - chunk start (int32, increasing
// per thread); lift(loc, lo) -> Acc.  Acc: init(), merge(l, r) [commutative+associative],
// finish(acc, plan, out).  kPrefix: the reducer needs the prodover early-exit second phase.

// sumover / dsumover: tmp += a over good elements; no good element in bad mode -> BAD (Ufunc.pd:102-110)
template <class T, class O> struct RSum {
  static constexpr bool kPrefix = false;
  static constexpr bool kRescan = false;
  static constexpr int kUnroll = (sizeof(O) <= 4) ? 4 : 8;
  struct Loc { O s; int32_t any; };
  struct Acc { O s; int32_t any; int32_t pad; };
  static __device__ __forceinline__ Loc linit() { Loc x; x.s = O(0); x.any = 0; return x; }
  static __device__ __forceinline__ void lpush(Loc &x, T v, int32_t) { x.s = wrap_add<O>(x.s, (O)v); x.any = 1; }
  // 64-bit integer sums in BAD mode: add 0 for a BAD element instead of branching around the add (the branch kept the
  // compiler from holding the trip's loads in flight: 0.80 of peak)
  static constexpr bool kPushSel = tt<O>::is_int && sizeof(O) == 8;
  static __device__ __forceinline__ void lpush_sel(Loc &x, T v, int32_t, bool bad) {
    x.s = wrap_add<O>(x.s, bad ? O(0) : (O)v); x.any |= (int32_t)!bad;
  }
  // 8/16-bit integers into a 32-bit sum: a whole 16-byte image at a time (see swar_* above)
  static constexpr bool kPack = tt<T>::is_int && sizeof(T) <= 2 && std::is_same<O, int32_t>::value;
  template <int BADK> static __device__ __forceinline__ void lpush_pack(Loc &x, const Pack<T> &r, T abad) {
    const uint32_t badw = swar_splat<T>(abad);
    const uint32_t w[4] = {r.q.x, r.q.y, r.q.z, r.q.w};
    int32_t nbad = 0, s = (int32_t)x.s;
    // BAD lanes are summed like the others and taken out again as badvalue x (number of BAD lanes): no mask
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if constexpr (BADK == 1) nbad += swar_eq_count<T>(w[i], badw);
      s = swar_sum<T>(w[i], s);
    }
    if constexpr (BADK == 1) s -= (int32_t)abad * nbad;
    x.s = (O)s;
    x.any |= (nbad != (int32_t)(16 / sizeof(T)));
  }
  static __device__ __forceinline__ Acc lift(const Loc &l, int64_t) { Acc x; x.s = l.s; x.any = l.any; x.pad = 0; return x; }
  static __device__ __forceinline__ Acc init() { Acc x; x.s = O(0); x.any = 0; x.pad = 0; return x; }
  static __device__ __forceinline__ Acc merge(const Acc &l, const Acc &r) { Acc x; x.s = wrap_add<O>(l.s, r.s); x.any = l.any | r.any; x.pad = 0; return x; }
  static __device__ __forceinline__ void finish(const Acc &x, const RdPlan &p, O *out) {
    *out = (p.badmode && !x.any) ? from_bits<O>(p.bbad) : x.s;
  }
};
// prodover / dprodover (Ufunc.pd:91,102-110): `tmp *= a; if (tmp == 0) break;`.  For integer
// outputs 0 absorbs everything, so the plain wrapped product is already the answer.  For float
// outputs the loop stops at the first zero element z, so the answer is prod(a[0..z]) — the sign of
// the zero comes from the prefix only, and an inf/NaN prefix gives NaN (inf*0) which never compares
// equal to 0, so the loop runs on and stays NaN.  z is reduced here (min); the kernel then runs a
// second pass over [0, z) for rows that have one (kPrefix).
// The loop ALSO stops when the running product underflows to zero before any zero element ([1e-200, 1e-200, inf] is 0
// in the reference; a product taken in another order is NaN, or finite).  That depends on the sequential order, so the
// reducer only detects that it CAN happen, in two steps.  Hot loop (3 integer instructions per element): the smallest
// non-zero magnitude of the row; n * min(0, floor(log2 of it)) bounds every prefix product from below, and rows of
// factors >= 1 never get past this test.  Rows that do are summed again by the group with
// negl = sum of min(0, log2|a|) (RNegLog, one MUFU per element): 2^negl is the tight bound.  Only rows whose tight
// bound reaches the subnormal range are recomputed by one thread in the reference's order (rd_prod_sequential) —
// which stops early exactly where the reference does.
template <class T> __device__ __forceinline__ float neg_log2_abs(T v) {
  if constexpr (tt<T>::is_int) return 0.0f;
  else if constexpr (sizeof(T) == 4) {
    const float a = fabsf(v);
    return (a < 1.0f && a != 0.0f) ? __log2f(a) : 0.0f;
  } else {
    const double a = fabs(v);
    if (!(a < 1.0) || a == 0.0) return 0.0f;
    const int hi = __double2hiint(a), e = (hi >> 20) & 0x7ff;
    if (e == 0) return -1100.0f;                                     // subnormal: straight into the recomputed class
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(a));   // [1, 2)
    return (float)(e - 1023) + __log2f((float)m);
  }
}
// magnitude key of a float/double: monotonic in |v| over the non-zero values, zero -> 0xffffffff (never the minimum)
template <class T> __device__ __forceinline__ uint32_t mag_key(T v, bool is_zero) {
  if constexpr (tt<T>::is_int) return 0xffffffffu;
  else if constexpr (sizeof(T) == 4) return (__float_as_uint(v) & 0x7fffffffu) - (uint32_t)is_zero;
  else return ((uint32_t)__double2hiint(v) & 0x7fffffffu) - (uint32_t)is_zero;   // a subnormal below 2^-1042: key 0, the smallest
}
template <class T, class O> struct RProd {
  static constexpr bool kPrefix = !tt<O>::is_int;
  static constexpr bool kUnderflow = kPrefix && !tt<T>::is_int;      // integer factors are 0 or >= 1 in magnitude
  static constexpr bool kRescan = false;
  static constexpr int kUnroll = (sizeof(O) <= 4 && tt<O>::is_int) ? 4 : 8;
  struct Loc { O s; int32_t any; int32_t z; uint32_t mkey; };
  struct Acc { O s; int32_t any; uint32_t mkey; int64_t z; };
  static __device__ __forceinline__ Loc linit() { Loc x; x.s = O(1); x.any = 0; x.z = 0x7fffffff; x.mkey = 0xffffffffu; return x; }
  static __device__ __forceinline__ void lpush(Loc &x, T v, int32_t rel) {
    x.s = wrap_mul<O>(x.s, (O)v); x.any = 1;
    if constexpr (kPrefix) {
      const bool zero = (O)v == O(0);
      if (zero && rel < x.z) x.z = rel;
      if constexpr (kUnderflow) { const uint32_t k = mag_key<T>(v, zero); x.mkey = k < x.mkey ? k : x.mkey; }
    }
  }
  // BAD mode without a branch: a BAD element is the factor 1 (never zero, never the smallest magnitude below 1)
  static constexpr bool kPushSel = kUnderflow;
  static __device__ __forceinline__ void lpush_sel(Loc &x, T v, int32_t rel, bool bad) {
    const T w = bad ? T(1) : v;
    const int32_t any = x.any | (int32_t)!bad;
    lpush(x, w, rel);
    x.any = any;
  }
  // binades below which a product is zero, less a margin for the approximate logarithms of the second step
  static constexpr float kZeroLog2 = sizeof(O) == 4 ? -137.0f : -1060.0f;
  // step 1: can n factors no smaller than the row's smallest one underflow?
  static __device__ __forceinline__ bool may_underflow(const Acc &x, int64_t n) {
    if constexpr (!kUnderflow) return false;
    else {
      if (x.mkey == 0xffffffffu) return false;                      // no non-zero factor at all
      const int e = sizeof(T) == 4 ? (int)(x.mkey >> 23) - 127 : (int)(x.mkey >> 20) - 1023;   // floor(log2 min|a|)
      return e < 0 && (float)n * (float)e <= kZeroLog2;
    }
  }
  // 8/16-bit integers into a 32-bit product: BAD lanes are replaced by 1 on whole words (no per-element compare /
  // branch), then one multiply per lane
  static constexpr bool kPack = tt<T>::is_int && sizeof(T) <= 2 && std::is_same<O, int32_t>::value;
  template <int BADK> static __device__ __forceinline__ void lpush_pack(Loc &x, const Pack<T> &r, T abad) {
    const uint32_t badw = swar_splat<T>(abad), onew = swar_splat<T>(T(1));
    const uint32_t w[4] = {r.q.x, r.q.y, r.q.z, r.q.w};
    int32_t nbad = 0;
    uint32_t s = (uint32_t)x.s;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint32_t g = w[i];
      if constexpr (BADK == 1) { const uint32_t m = swar_eq_mask<T>(g, badw, nbad); g = (g & ~m) | (onew & m); }
#pragma unroll
      for (int k = 0; k < (int)(4 / sizeof(T)); k++) s *= (uint32_t)(int32_t)(T)(g >> (8 * sizeof(T) * k));
    }
    x.s = (O)s;
    x.any |= (nbad != (int32_t)(16 / sizeof(T)));
  }
  static __device__ __forceinline__ Acc lift(const Loc &l, int64_t lo) {
    Acc x; x.s = l.s; x.any = l.any; x.mkey = l.mkey; x.z = (l.z == 0x7fffffff) ? RD_NOIDX : lo + l.z; return x;
  }
  static __device__ __forceinline__ Acc init() { Acc x; x.s = O(1); x.any = 0; x.mkey = 0xffffffffu; x.z = RD_NOIDX; return x; }
  static __device__ __forceinline__ Acc merge(const Acc &l, const Acc &r) {
    Acc x; x.s = wrap_mul<O>(l.s, r.s); x.any = l.any | r.any; x.mkey = l.mkey < r.mkey ? l.mkey : r.mkey; x.z = l.z < r.z ? l.z : r.z; return x;
  }
  static __device__ __forceinline__ void finish(const Acc &x, const RdPlan &p, O *out) {
    *out = (p.badmode && !x.any) ? from_bits<O>(p.bbad) : x.s;
  }
};
// average / daverage (Ufunc.pd:417-430): tmp / cnt evaluated with C's usual arithmetic
// conversions (cnt is PDL_Indx = int64); cnt == 0 -> BAD (bad mode) or 0 / NaN (good mode).
template <class T, class O> struct RAvg {
  static constexpr bool kPrefix = false;
  static constexpr bool kRescan = false;
  static constexpr int kUnroll = (sizeof(O) <= 4) ? 4 : 8;
  struct Loc { O s; int32_t cnt; };
  struct Acc { O s; int64_t cnt; };
  static __device__ __forceinline__ Loc linit() { Loc x; x.s = O(0); x.cnt = 0; return x; }
  static __device__ __forceinline__ void lpush(Loc &x, T v, int32_t) { x.s = wrap_add<O>(x.s, (O)v); x.cnt++; }
  static constexpr bool kPushSel = tt<O>::is_int && sizeof(O) == 8;
  static __device__ __forceinline__ void lpush_sel(Loc &x, T v, int32_t, bool bad) {
    x.s = wrap_add<O>(x.s, bad ? O(0) : (O)v); x.cnt += (int32_t)!bad;
  }
  static constexpr bool kPack = tt<T>::is_int && sizeof(T) <= 2 && std::is_same<O, int32_t>::value;
  template <int BADK> static __device__ __forceinline__ void lpush_pack(Loc &x, const Pack<T> &r, T abad) {
    const uint32_t badw = swar_splat<T>(abad);
    const uint32_t w[4] = {r.q.x, r.q.y, r.q.z, r.q.w};
    int32_t nbad = 0, s = (int32_t)x.s;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if constexpr (BADK == 1) nbad += swar_eq_count<T>(w[i], badw);
      s = swar_sum<T>(w[i], s);
    }
    if constexpr (BADK == 1) s -= (int32_t)abad * nbad;
    x.s = (O)s;
    x.cnt += (int32_t)(16 / sizeof(T)) - nbad;
  }
  static __device__ __forceinline__ Acc lift(const Loc &l, int64_t) { Acc x; x.s = l.s; x.cnt = l.cnt; return x; }
  static __device__ __forceinline__ Acc init() { Acc x; x.s = O(0); x.cnt = 0; return x; }
  static __device__ __forceinline__ Acc merge(const Acc &l, const Acc &r) { Acc x; x.s = wrap_add<O>(l.s, r.s); x.cnt = l.cnt + r.cnt; return x; }
  static __device__ __forceinline__ void finish(const Acc &x, const RdPlan &p, O *out) {
    if (x.cnt == 0) {
      if (p.badmode) *out = from_bits<O>(p.bbad);
      else if constexpr (tt<O>::is_int) *out = O(0);
      else *out = nan_of<O>();
      return;
    }
    if constexpr (!tt<O>::is_int) *out = x.s / (O)x.cnt;
    else if constexpr (sizeof(O) == 8 && tt<O>::is_uns) *out = (O)((uint64_t)x.s / (uint64_t)x.cnt);
    else *out = (O)((int64_t)x.s / x.cnt);
  }
};
// minimum / maximum / _ind (Ufunc.pd:455-465,481-491).  Sequential rule: cur is replaced when
// (a OP cur) or cur is NaN (or nothing was taken yet).  Loc applies exactly that rule, with "nothing
// yet" encoded as idx < 0 (and cur = NaN for float types so one test covers both).  Closed form used
// by the merge: the first-in-index-order extreme of the non-NaN good values; if every good value is
// NaN, the LAST NaN; if there is no good value, BAD.
template <class T, class O, bool ISMAX, bool WANT_IND> struct RMinMaxExact {
  static constexpr bool kPrefix = false;
  static constexpr bool kRescan = false;
  static constexpr int kUnroll = 8;
  struct Loc { T cur; int32_t idx; };
  struct Acc { T cur; int64_t idx; int32_t state; int32_t pad; };  // state 0 empty, 1 non-NaN, 2 NaN only
  static __device__ __forceinline__ Loc linit() {
    Loc x; x.idx = -1;
    if constexpr (tt<T>::is_int) x.cur = T(0); else x.cur = nan_of<T>();
    return x;
  }
  static __device__ __forceinline__ void lpush(Loc &x, T v, int32_t rel) {
    bool take;
    if constexpr (tt<T>::is_int) take = (ISMAX ? (v > x.cur) : (v < x.cur)) || (x.idx < 0);
    else take = (ISMAX ? (v > x.cur) : (v < x.cur)) || (x.cur != x.cur);
    x.cur = take ? v : x.cur;
    x.idx = take ? rel : x.idx;
  }
  static __device__ __forceinline__ Acc lift(const Loc &l, int64_t lo) {
    Acc x; x.cur = l.cur; x.idx = l.idx < 0 ? -1 : lo + l.idx; x.pad = 0;
    x.state = l.idx < 0 ? 0 : (t_isnan(l.cur) ? 2 : 1);
    return x;
  }
  static __device__ __forceinline__ Acc init() { Acc x; x.cur = T(0); x.idx = -1; x.state = 0; x.pad = 0; return x; }
  static __device__ __forceinline__ bool better(T v, int64_t i, T cur, int64_t idx) {
    return (ISMAX ? (v > cur) : (v < cur)) || (v == cur && i < idx);
  }
  static __device__ __forceinline__ Acc merge(const Acc &l, const Acc &r) {
    if (r.state == 0) return l;
    if (l.state == 0) return r;
    if (l.state == 1 && r.state == 1) return better(r.cur, r.idx, l.cur, l.idx) ? r : l;
    if (l.state == 1) return l;
    if (r.state == 1) return r;
    return (r.idx > l.idx) ? r : l;
  }
  static __device__ __forceinline__ void finish(const Acc &x, const RdPlan &p, O *out) {
    if (x.state == 0) *out = from_bits<O>(p.bbad);
    else if constexpr (WANT_IND) *out = (O)x.idx;
    else *out = (O)x.cur;
  }
};
// Hot-loop version of the same reducer: cur starts at the identity (+inf / type max for minimum,
// -inf / type min for maximum) and an element is taken iff it is STRICTLY better — two instructions
// fewer per element than the exact rule (no "cur is NaN / nothing taken yet" test), NaNs are never
// taken.  What it cannot represent is a thread whose good elements are all equal to the identity or
// NaN: then nothing was taken (idx < 0) and the kernel re-walks just that thread's elements with
// RMinMaxExact (kRescan).  Acc / merge / finish are the exact reducer's, so results are identical.
template <class T, class O, bool ISMAX, bool WANT_IND> struct RMinMax {
  using Exact = RMinMaxExact<T, O, ISMAX, WANT_IND>;
  using Acc = typename Exact::Acc;
  static constexpr bool kPrefix = false;
  static constexpr bool kRescan = true;
  static constexpr int kUnroll = 8;
  struct Loc { T cur; int32_t idx; };
  static __device__ __forceinline__ T identity() {
    if constexpr (sizeof(T) == 4 && !tt<T>::is_int) return __uint_as_float(ISMAX ? 0xff800000u : 0x7f800000u);
    else if constexpr (!tt<T>::is_int) return __longlong_as_double(ISMAX ? (long long)0xfff0000000000000ull : 0x7ff0000000000000ll);
    else if constexpr (tt<T>::is_uns) return ISMAX ? T(0) : T(~T(0));
    else { using U = typename std::make_unsigned<T>::type; const T mx = (T)(U(~U(0)) >> 1); return ISMAX ? (T)(-mx - 1) : mx; }
  }
  static __device__ __forceinline__ Loc linit() { Loc x; x.cur = identity(); x.idx = -1; return x; }
  static __device__ __forceinline__ void lpush(Loc &x, T v, int32_t rel) {
    const bool take = ISMAX ? (v > x.cur) : (v < x.cur);
    x.cur = take ? v : x.cur;
    x.idx = take ? rel : x.idx;
  }
  static __device__ __forceinline__ bool needs_rescan(const Loc &l) { return l.idx < 0; }
  static __device__ __forceinline__ Acc lift(const Loc &l, int64_t lo) {
    Acc x; x.cur = l.cur; x.idx = lo + l.idx; x.state = 1; x.pad = 0; return x;
  }
  static __device__ __forceinline__ Acc init() { return Exact::init(); }
  static __device__ __forceinline__ Acc merge(const Acc &l, const Acc &r) { return Exact::merge(l, r); }
  static __device__ __forceinline__ void finish(const Acc &x, const RdPlan &p, O *out) { Exact::finish(x, p, out); }
};

// Integer minimum / maximum WITHOUT index output: equal integers are bit-identical, so "first wins" needs no
// index at all — the hot loop is one IMNMX per element (plus the BAD select), which is what the 8/16-bit types
// need to stay HBM-bound (16 elements per 128-bit load).
template <class T, bool ISMAX> struct RMinMaxInt {
  static constexpr bool kPrefix = false;
  static constexpr bool kRescan = false;
  static constexpr int kUnroll = 4;
  // 8/16-bit types: `pk` holds 4 (2) independent running extremes, one per lane of a 32-bit word, updated by
  // one packed min/max per word (VIMNMX.x16x2; a few LOP3/PRMT for bytes); BAD lanes are first overwritten with
  // the identity.  lift() folds the lanes into `cur`.
  static constexpr bool kPack = sizeof(T) <= 2;
  struct Loc { T cur; int32_t any; uint32_t pk; };
  struct Acc { T cur; int32_t any; int32_t pad; int32_t pad2; };
  static __device__ __forceinline__ T identity() {
    if constexpr (tt<T>::is_uns) return ISMAX ? T(0) : T(~T(0));
    else { using U = typename std::make_unsigned<T>::type; const T mx = (T)(U(~U(0)) >> 1); return ISMAX ? (T)(-mx - 1) : mx; }
  }
  // the OTHER extreme of the type.  When the badvalue is that one (minimum of a signed row, maximum of an unsigned one,
  // with the default badvalues), adding (max) or subtracting (min) one in every lane, wrapping, turns exactly the BAD
  // lanes into the identity and keeps the order of all others: again no mask (BADK 4); lift() undoes the shift.
  static __device__ __forceinline__ T anti_identity() {
    if constexpr (tt<T>::is_uns) return ISMAX ? T(~T(0)) : T(0);
    else { using U = typename std::make_unsigned<T>::type; const T mx = (T)(U(~U(0)) >> 1); return ISMAX ? mx : (T)(-mx - 1); }
  }
  static __device__ __forceinline__ uint32_t shift_packed(uint32_t w) {
    if constexpr (sizeof(T) == 1) return ISMAX ? __vadd4(w, 0x01010101u) : __vsub4(w, 0x01010101u);
    else return ISMAX ? __vadd2(w, 0x00010001u) : __vsub2(w, 0x00010001u);
  }
  static __device__ __forceinline__ T pick(T a, T b) { return ISMAX ? (b > a ? b : a) : (b < a ? b : a); }
  static __device__ __forceinline__ uint32_t pick_packed(uint32_t a, uint32_t b) {
    if constexpr (sizeof(T) == 1) {
      if constexpr (tt<T>::is_uns) return ISMAX ? __vmaxu4(a, b) : __vminu4(a, b); else return ISMAX ? __vmaxs4(a, b) : __vmins4(a, b);
    } else {
      if constexpr (tt<T>::is_uns) return ISMAX ? __vmaxu2(a, b) : __vminu2(a, b); else return ISMAX ? __vmaxs2(a, b) : __vmins2(a, b);
    }
  }
  static __device__ __forceinline__ Loc linit() {
    Loc x; x.cur = identity(); x.any = 0; x.pk = 0;
    if constexpr (kPack) x.pk = swar_splat<T>(identity());
    return x;
  }
  static __device__ __forceinline__ void lpush(Loc &x, T v, int32_t) { x.cur = pick(x.cur, v); x.any |= 1; }
  template <int BADK> static __device__ __forceinline__ void lpush_pack(Loc &x, const Pack<T> &r, T abad) {
    const uint32_t badw = swar_splat<T>(abad), identw = swar_splat<T>(identity());
    const uint32_t w[4] = {r.q.x, r.q.y, r.q.z, r.q.w};
    int32_t nbad = 0;
#pragma unroll
    if constexpr (BADK == 3) {
      // the badvalue IS this reducer's identity (the default badvalues are the type extremes: minimum of an unsigned
      // row, maximum of a signed one): BAD lanes cannot win, so no mask — only "was there a good lane at all"
      int32_t good = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) { x.pk = pick_packed(x.pk, w[i]); good |= (w[i] != badw); }
      x.any |= good;
    } else if constexpr (BADK == 4) {
      int32_t good = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) { x.pk = pick_packed(x.pk, shift_packed(w[i])); good |= (w[i] != badw); }
      x.any |= good | 2;                       // bit 1: the lanes of pk are in the shifted domain
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        uint32_t g = w[i];
        if constexpr (BADK == 1) { const uint32_t m = swar_eq_mask<T>(g, badw, nbad); g = (g & ~m) | (identw & m); }
        x.pk = pick_packed(x.pk, g);
      }
      x.any |= (nbad != (int32_t)(16 / sizeof(T)));
    }
  }
  static constexpr bool kBadIdentity = kPack;   // rd_row may pick BADK 3 when the badvalue equals identity()
  static __device__ __forceinline__ Acc lift(const Loc &l, int64_t) {
    Acc x; x.cur = l.cur; x.any = l.any & 1; x.pad = 0; x.pad2 = 0;
    if constexpr (kPack) {
      const bool shifted = (l.any & 2) != 0;
#pragma unroll
      for (int k = 0; k < (int)(4 / sizeof(T)); k++) {
        T v = (T)(l.pk >> (8 * sizeof(T) * k));
        if (shifted) {
          if (v == identity()) continue;       // only BAD lanes (or none at all) end up on the identity
          v = ISMAX ? (T)(v - 1) : (T)(v + 1);
        }
        x.cur = pick(x.cur, v);
      }
    }
    return x;
  }
  static __device__ __forceinline__ Acc init() { Acc x; x.cur = identity(); x.any = 0; x.pad = 0; x.pad2 = 0; return x; }
  static __device__ __forceinline__ Acc merge(const Acc &l, const Acc &r) {
    Acc x; x.any = l.any | r.any; x.pad = 0; x.pad2 = 0;
    x.cur = pick(l.cur, r.cur);
    return x;
  }
  static __device__ __forceinline__ void finish(const Acc &x, const RdPlan &p, T *out) {
    *out = x.any ? x.cur : from_bits<T>(p.bbad);
  }
};

// minimum_ind / maximum_ind of 8/16-bit integer rows in TWO phases: the packed value reduction above (one
// VIMNMX-class instruction per word), then a search for the FIRST element equal to the extreme (one packed
// equality test per word, the row comes back from L1/L2).  Integers have no NaN and no -0, so "first index of the
// extreme value" is exactly the reference's strict-compare rule (Ufunc.pd:481-491).  kFindFirst makes the kernel
// run phase 2; rows that are cut into chunks (nchunks > 1) use the per-element reducer instead.
template <class T, bool ISMAX> struct RMinMaxIntInd : RMinMaxInt<T, ISMAX> {
  static constexpr bool kFindFirst = true;
  using Acc = typename RMinMaxInt<T, ISMAX>::Acc;
  // only reached by the (never launched) chunked instantiations: the index comes from phase 2 in the row kernel
  static __device__ __forceinline__ void finish(const Acc &, const RdPlan &p, int64_t *out) { *out = from_bits<int64_t>(p.bbad); }
};
template <class R, class = void> struct rd_kfind { static constexpr bool value = false; };
template <class R> struct rd_kfind<R, std::void_t<decltype(R::kFindFirst)>> { static constexpr bool value = R::kFindFirst; };

// andover orover zcover xorover (logical) and bandover borover bxorover (bitwise), Ufunc.pd:143-187.
// KIND: 0 and, 1 or, 2 zc, 3 xor, 4 band, 5 bor, 6 bxor.  Output type == input type.
template <class T, int KIND> struct RBits {
  static constexpr bool kPrefix = false;
  static constexpr bool kRescan = false;
  static constexpr int kUnroll = (sizeof(T) <= 4) ? 4 : 8;
  using U = typename tt<T>::wide_u;
  struct Acc { U v; int32_t any; };
  using Loc = Acc;
  static __device__ __forceinline__ Acc init() {
    Acc x; x.any = 0;
    x.v = (KIND == 0 || KIND == 2) ? U(1) : (KIND == 4) ? ~U(0) : U(0);
    return x;
  }
  static __device__ __forceinline__ Loc linit() { return init(); }
  static __device__ __forceinline__ Acc lift(const Loc &l, int64_t) { return l; }
  static __device__ __forceinline__ U bits(T a) { if constexpr (tt<T>::is_int) return (U)a; else return U(0); }
  static __device__ __forceinline__ void lpush(Acc &x, T a, int32_t) {
    x.any = 1;
    if constexpr (KIND == 0) x.v &= U(a != 0);
    else if constexpr (KIND == 1) x.v |= U(a != 0);
    else if constexpr (KIND == 2) x.v &= U(a == 0);
    else if constexpr (KIND == 3) x.v ^= U(a != 0);
    else if constexpr (KIND == 4) x.v &= bits(a);
    else if constexpr (KIND == 5) x.v |= bits(a);
    else x.v ^= bits(a);
  }
  // 8/16-bit integers: whole 16-byte images at a time.  Bitwise kinds fold the four words (BAD lanes first set to
  // the kind's identity) and then the lanes; logical kinds only need "is any good lane (non)zero" per word.
  static constexpr bool kPack = tt<T>::is_int && sizeof(T) <= 2;
  template <int BADK> static __device__ __forceinline__ void lpush_pack(Acc &x, const Pack<T> &r, T abad) {
    constexpr uint32_t LANE = sizeof(T) == 1 ? 0xffu : 0xffffu;
    const uint32_t badw = swar_splat<T>(abad);
    const uint32_t w[4] = {r.q.x, r.q.y, r.q.z, r.q.w};
    int32_t nbad = 0, nzero = 0;
    uint32_t fold = (KIND == 4) ? 0xffffffffu : 0u;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint32_t m = 0;
      if constexpr (BADK == 1) m = swar_eq_mask<T>(w[i], badw, nbad);
      if constexpr (KIND == 4) fold &= (w[i] | m);
      else if constexpr (KIND == 5) fold |= (w[i] & ~m);
      else if constexpr (KIND == 6) fold ^= (w[i] & ~m);
      else if constexpr (KIND == 1 || KIND == 2) fold |= (w[i] & ~m);            // any good lane non-zero?
      else {                                                                       // KIND 0 / 3: count the good zero lanes
        int32_t nz_all = 0;
        const uint32_t z = swar_eq_mask<T>(w[i], 0u, nz_all);
        if constexpr (BADK == 1) nzero += __popc(z & ~m) / (int)(8 * sizeof(T)); else nzero += nz_all;
      }
    }
    const int32_t ngood = (int32_t)(16 / sizeof(T)) - nbad;
    x.any |= (ngood != 0);
    if constexpr (KIND == 4) { fold &= fold >> 16; if constexpr (sizeof(T) == 1) fold &= fold >> 8; x.v &= (fold | ~LANE); }
    else if constexpr (KIND == 5) { fold |= fold >> 16; if constexpr (sizeof(T) == 1) fold |= fold >> 8; x.v |= (fold & LANE); }
    else if constexpr (KIND == 6) { fold ^= fold >> 16; if constexpr (sizeof(T) == 1) fold ^= fold >> 8; x.v ^= (fold & LANE); }
    else if constexpr (KIND == 1) x.v |= U(fold != 0);
    else if constexpr (KIND == 2) x.v &= U(fold == 0);
    else if constexpr (KIND == 0) x.v &= U(nzero == 0);
    else x.v ^= U((ngood - nzero) & 1);
  }
  static __device__ __forceinline__ Acc merge(const Acc &l, const Acc &r) {
    Acc x; x.any = l.any | r.any;
    if constexpr (KIND == 0 || KIND == 2 || KIND == 4) x.v = l.v & r.v;
    else if constexpr (KIND == 1 || KIND == 5) x.v = l.v | r.v;
    else x.v = l.v ^ r.v;
    return x;
  }
  static __device__ __forceinline__ void finish(const Acc &x, const RdPlan &p, T *out) {
    *out = (p.badmode && !x.any) ? from_bits<T>(p.bbad) : (T)x.v;
  }
};

// nbadover / ngoodover (lib/PDL/Bad.pd:418-480): counts, output indx.  Only good elements reach
// lpush, so the accumulator counts those; nbad = n - ngood (good mode: nbad = 0, ngood = n).
template <class T, bool GOOD> struct RCount {
  static constexpr bool kPrefix = false;
  static constexpr bool kRescan = false;
  static constexpr int kUnroll = 4;
  struct Loc { int32_t cnt; };
  struct Acc { int64_t cnt; };
  static __device__ __forceinline__ Loc linit() { Loc x; x.cnt = 0; return x; }
  static __device__ __forceinline__ void lpush(Loc &x, T, int32_t) { x.cnt++; }
  static __device__ __forceinline__ Acc lift(const Loc &l, int64_t) { Acc x; x.cnt = l.cnt; return x; }
  static __device__ __forceinline__ Acc init() { Acc x; x.cnt = 0; return x; }
  static __device__ __forceinline__ Acc merge(const Acc &l, const Acc &r) { Acc x; x.cnt = l.cnt + r.cnt; return x; }
  static __device__ __forceinline__ void finish(const Acc &x, const RdPlan &p, int64_t *out) {
    *out = GOOD ? x.cnt : (p.n - x.cnt);
  }
};

// ---- row walk -------------------------------------------------------------------
// BADK: 0 = no BAD test (good-mode code path), 1 = BAD iff v == badvalue, 2 = BAD iff v is NaN,
// 3 / 4 = the badvalue is the packed min/max reducer's identity / the other extreme of the type (no mask needed)
// (per-ndarray NaN badvalue).  Hoisted to a template so the hot loop pays one compare at most.
template <class R, class = void> struct rd_kbadident { static constexpr bool value = false; };
template <class R> struct rd_kbadident<R, std::void_t<decltype(R::kBadIdentity)>> { static constexpr bool value = R::kBadIdentity; };

// reducers whose per-element body is long enough that skipping it with a branch costs more than running it on a
// neutral element: `lpush_sel(loc, v, rel, bad)`
template <class R, class = void> struct rd_kpushsel : std::false_type {};
template <class R> struct rd_kpushsel<R, std::void_t<decltype(R::kPushSel)>> : std::integral_constant<bool, R::kPushSel> {};

template <class R, class T, int BADK>
__device__ __forceinline__ void rd_push(typename R::Loc &loc, T v, int32_t rel, T abad) {
  if constexpr (rd_kpushsel<R>::value && (BADK == 1 || BADK == 2)) {
    R::lpush_sel(loc, v, rel, BADK == 1 ? (v == abad) : t_isnan(v));
    return;
  }
  if constexpr (BADK == 1 || BADK == 3 || BADK == 4) { if (v == abad) return; }
  if constexpr (BADK == 2) { if (t_isnan(v)) return; }
  R::lpush(loc, v, rel);
}

// Accumulate elements [lo, hi) of one row (hi - lo < 2^31); `lane` of `width` cooperating threads.
// Each thread visits its elements in increasing index order.
template <class R, class T, int BADK>
__device__ __forceinline__ void rd_row_k(typename R::Loc &loc, const T *row, int64_t lo, int64_t hi, int64_t inc,
                                         int lane, int width, T abad) {
  constexpr int VEC = 16 / sizeof(T);
  const T *base = row + lo * inc;         // element `rel` lives at base[rel * inc]
  const int32_t len = (int32_t)(hi - lo);
  if (inc == 1) {
    // peel to 16-byte alignment, then 128-bit loads with R::kUnroll in flight, then the tail
    const uintptr_t addr = (uintptr_t)base;
    int32_t head = (int32_t)(((16 - (addr & 15)) & 15) / sizeof(T));
    if (head > len) head = len;
    for (int32_t i = lane; i < head; i += width) rd_push<R, T, BADK>(loc, base[i], i, abad);
    const int32_t nv = (len - head) / VEC;
    const uint4 *vp = reinterpret_cast<const uint4 *>(base + head);
    int32_t j = lane;
    for (; j + (R::kUnroll - 1) * width < nv; j += R::kUnroll * width) {
      Pack<T> r[R::kUnroll];
#pragma unroll
      for (int u = 0; u < R::kUnroll; u++) r[u].q = vp[j + u * width];
#pragma unroll
      for (int u = 0; u < R::kUnroll; u++) {
        if constexpr (rd_kpack<R>::value) R::template lpush_pack<BADK>(loc, r[u], abad);
        else {
          const int32_t e0 = head + (j + u * width) * VEC;
#pragma unroll
          for (int k = 0; k < VEC; k++) rd_push<R, T, BADK>(loc, r[u].e[k], e0 + k, abad);
        }
      }
    }
    for (; j < nv; j += width) {
      Pack<T> r; r.q = vp[j];
      if constexpr (rd_kpack<R>::value) R::template lpush_pack<BADK>(loc, r, abad);
      else {
        const int32_t e0 = head + j * VEC;
#pragma unroll
        for (int k = 0; k < VEC; k++) rd_push<R, T, BADK>(loc, r.e[k], e0 + k, abad);
      }
    }
    for (int32_t i = head + nv * VEC + lane; i < len; i += width) rd_push<R, T, BADK>(loc, base[i], i, abad);
  } else {
    int32_t i = lane;
    for (; i + (R::kUnroll - 1) * width < len; i += R::kUnroll * width) {
      T v[R::kUnroll];
#pragma unroll
      for (int u = 0; u < R::kUnroll; u++) v[u] = base[(int64_t)(i + u * width) * inc];
#pragma unroll
      for (int u = 0; u < R::kUnroll; u++) rd_push<R, T, BADK>(loc, v[u], i + u * width, abad);
    }
    for (; i < len; i += width) rd_push<R, T, BADK>(loc, base[(int64_t)i * inc], i, abad);
  }
}

template <class R, class T, bool BAD>
__device__ __forceinline__ void rd_row(typename R::Loc &loc, const T *row, int64_t lo, int64_t hi, int64_t inc,
                                       int lane, int width, T abad, bool abadnan) {
  if constexpr (!BAD) rd_row_k<R, T, 0>(loc, row, lo, hi, inc, lane, width, abad);
  else if constexpr (rd_kbadident<R>::value) {
    if (abad == R::identity()) rd_row_k<R, T, 3>(loc, row, lo, hi, inc, lane, width, abad);
    else if (abad == R::anti_identity()) rd_row_k<R, T, 4>(loc, row, lo, hi, inc, lane, width, abad);
    else rd_row_k<R, T, 1>(loc, row, lo, hi, inc, lane, width, abad);
  }
  else if constexpr (tt<T>::is_int) rd_row_k<R, T, 1>(loc, row, lo, hi, inc, lane, width, abad);
  else { if (abadnan) rd_row_k<R, T, 2>(loc, row, lo, hi, inc, lane, width, abad);
         else rd_row_k<R, T, 1>(loc, row, lo, hi, inc, lane, width, abad); }
}

// Phase 2 of the two-phase _ind reducers: a candidate for the smallest index in [0, n) with row[i] == e, such that
// the minimum of the candidates over the cooperating group IS that index (RD_NOIDX: this thread has none).
// Each warp of the group scans its own contiguous share of the row front to back, 32 vectors at a time, and
// stops at the first wavefront that contains a hit (`__any_sync`): with ties (the usual case for small integer
// types) the scan ends after a few hundred bytes; without them the row is re-read once from L1/L2.
template <class T>
__device__ __forceinline__ int64_t rd_first_eq(const T *row, int64_t n, int64_t inc, int lane, int width, T e) {
  constexpr int VEC = 16 / sizeof(T);
  if (width == 1) {                               // thread per row: plain sequential search
    for (int64_t i = 0; i < n; i++) if (row[i * inc] == e) return i;
    return RD_NOIDX;
  }
  const int nw = width >> 5, wid = lane >> 5, l = lane & 31;
  int64_t best = RD_NOIDX;
  if (inc == 1) {
    const uintptr_t addr = (uintptr_t)row;
    int64_t head = (int64_t)(((16 - (addr & 15)) & 15) / sizeof(T));
    if (head > n) head = n;
    if (wid == 0 && l < head && row[l] == e) best = l;                     // head < 16 elements
    const int64_t nv = (n - head) / VEC;
    const uint4 *vp = reinterpret_cast<const uint4 *>(row + head);
    const uint32_t ew = swar_splat<T>(e);
    const int64_t per = (nv + nw - 1) / nw;
    const int64_t j1 = (wid + 1) * per < nv ? (wid + 1) * per : nv;
    for (int64_t j0 = wid * per; j0 < j1; j0 += 32) {
      const int64_t j = j0 + l;
      int64_t cand = RD_NOIDX;
      if (j < j1) {
        const uint4 q = vp[j];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        int32_t dummy = 0;
#pragma unroll
        for (int i = 3; i >= 0; i--) {
          const uint32_t m = swar_eq_mask<T>(w[i], ew, dummy);
          if (m) cand = head + j * VEC + i * (int)(4 / sizeof(T)) + (__ffs(m) - 1) / (int)(8 * sizeof(T));
        }
      }
      if (__any_sync(0xffffffffu, cand != RD_NOIDX)) { best = cand < best ? cand : best; break; }
    }
    if (wid == nw - 1) {                                                   // tail < VEC elements
      const int64_t i = head + nv * VEC + l;
      if (i < n && row[i] == e && i < best) best = i;
    }
  } else {
    const int64_t per = (n + nw - 1) / nw;
    const int64_t i1 = (wid + 1) * per < n ? (wid + 1) * per : n;
    for (int64_t i0 = wid * per; i0 < i1; i0 += 32) {
      const int64_t i = i0 + l;
      const bool hit = i < i1 && row[i * inc] == e;
      if (__any_sync(0xffffffffu, hit)) { if (hit) best = i; break; }
    }
  }
  return best;
}
// minimum of an int64 over the cooperating group (MODE as in rd_group_reduce); smem: RD_THREADS/32 + 1 slots
template <int MODE>
__device__ __forceinline__ int64_t rd_group_min64(int64_t v, int64_t *smem) {
  if (MODE >= 1) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) { const int64_t o = __shfl_xor_sync(0xffffffffu, v, m); v = o < v ? o : v; }
  }
  if (MODE == 2) {
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) smem[w] = v;
    __syncthreads();
    if (w == 0) {
      v = (threadIdx.x < RD_THREADS / 32) ? smem[threadIdx.x] : RD_NOIDX;
#pragma unroll
      for (int m = (RD_THREADS / 64); m >= 1; m >>= 1) { const int64_t o = __shfl_xor_sync(0xffffffffu, v, m); v = o < v ? o : v; }
      if (threadIdx.x == 0) smem[RD_THREADS / 32] = v;
    }
    __syncthreads();
    v = smem[RD_THREADS / 32];
    __syncthreads();
  }
  return v;
}

__device__ __forceinline__ void rd_row_offsets(const RdPlan &p, int64_t row, int64_t &oa, int64_t &ob) {
  oa = 0; ob = 0;
  for (int d = 0; d < p.nd; d++) {
    int64_t q, i;
    if (d == p.nd - 1) { i = row; q = 0; }
    else if ((uint64_t)row <= 0xffffffffull && (uint64_t)p.dims[d] <= 0xffffffffull) {
      const uint32_t q32 = (uint32_t)row / (uint32_t)p.dims[d]; q = q32; i = (uint32_t)row - q32 * (uint32_t)p.dims[d];
    } else { q = row / p.dims[d]; i = row - q * p.dims[d]; }
    oa += i * p.sa[d]; ob += i * p.sb[d];
    row = q;
  }
}

// Reduce `acc` over the cooperating group; on return every thread of the group holds the total.
// MODE 2 uses `smem` (RD_THREADS/32 + 1 slots) and two barriers.
template <class R, int MODE>
__device__ __forceinline__ typename R::Acc rd_group_reduce(typename R::Acc acc, typename R::Acc *smem) {
  if (MODE >= 1) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) acc = R::merge(acc, shfl_xor_acc(acc, m));
  }
  if (MODE == 2) {
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) smem[w] = acc;
    __syncthreads();
    if (w == 0) {
      acc = (threadIdx.x < RD_THREADS / 32) ? smem[threadIdx.x] : R::init();
#pragma unroll
      for (int m = (RD_THREADS / 64); m >= 1; m >>= 1) acc = R::merge(acc, shfl_xor_acc(acc, m));
      if (threadIdx.x == 0) smem[RD_THREADS / 32] = acc;
    }
    __syncthreads();
    acc = smem[RD_THREADS / 32];
    __syncthreads();  // smem is reused by the next reduction
  }
  return acc;
}

// prodover in the reference's own order (Ufunc.pd:102-110), by ONE thread: for the rows RProd::may_underflow flags
template <class T, class O, bool BAD>
__device__ __noinline__ O rd_prod_sequential(const T *row, int64_t n, int64_t inc, T abad, bool abadnan, O bbad) {
  O tmp = O(1);
  bool flag = false;
  for (int64_t i = 0; i < n; i++) {
    const T v = row[i * inc];
    if (BAD && is_bad(v, abad, abadnan)) continue;
    flag = true;
    tmp = tmp * (O)v;
    if (tmp == O(0)) break;
  }
  return (BAD && !flag) ? bbad : tmp;
}

// step 2 of the underflow test: negl = sum of min(0, log2|a|) over the good non-zero elements of the row, by the group
template <class T> struct RNegLog {
  static constexpr bool kPrefix = false;
  static constexpr bool kRescan = false;
  static constexpr int kUnroll = 4;
  struct Acc { float s; int32_t any; };
  using Loc = Acc;
  static __device__ __forceinline__ Acc init() { Acc x; x.s = 0.0f; x.any = 0; return x; }
  static __device__ __forceinline__ Loc linit() { return init(); }
  static __device__ __forceinline__ void lpush(Loc &x, T v, int32_t) { x.s += neg_log2_abs<T>(v); }
  static __device__ __forceinline__ Acc lift(const Loc &l, int64_t) { return l; }
  static __device__ __forceinline__ Acc merge(const Acc &l, const Acc &r) { Acc x; x.s = l.s + r.s; x.any = 0; return x; }
};
template <class T, bool BAD, int MODE>
__device__ __forceinline__ float rd_neg_log(const T *row, int64_t n, int64_t inc, int lane, int width, T abad, bool abadnan, void *smem) {
  using P = RNegLog<T>;
  typename P::Acc tot = P::init();
  for (int64_t lo = 0; lo < n; lo += 0x40000000ll) {
    const int64_t hi = (lo + 0x40000000ll < n) ? lo + 0x40000000ll : n;
    typename P::Loc loc = P::linit();
    rd_row<P, T, BAD>(loc, row, lo, hi, inc, lane, width, abad, abadnan);
    tot = P::merge(tot, loc);
  }
  tot = rd_group_reduce<P, MODE>(tot, reinterpret_cast<typename P::Acc *>(smem));
  return tot.s;
}

// prodover second phase: the product of the good elements of [0, z) times a[z], by the group.
template <class R, class T, class O, bool BAD, int MODE>
__device__ __forceinline__ O rd_prod_prefix(const T *row, int64_t z, int64_t inc, int lane, int width,
                                            T abad, bool abadnan, typename RProd<T, O>::Acc *smem) {
  using P = RProd<T, O>;
  typename P::Acc tot = P::init();
  for (int64_t lo = 0; lo < z; lo += 0x40000000ll) {
    const int64_t hi = (lo + 0x40000000ll < z) ? lo + 0x40000000ll : z;
    typename P::Loc loc = P::linit();
    rd_row<P, T, BAD>(loc, row, lo, hi, inc, lane, width, abad, abadnan);
    tot = P::merge(tot, P::lift(loc, lo));
  }
  tot = rd_group_reduce<P, MODE>(tot, smem);
  return tot.s * (O)row[z * inc];
}

// MODE: 0 thread/row, 1 warp/row, 2 CTA/row.  blockIdx.y = chunk of n.
template <class R> constexpr int rd_min_blocks() { return R::kUnroll == 4 ? 8 : 5; }

// launch bounds: the light reducers (kUnroll 4) are held to 32 registers = 8 CTAs/SM; the rest to 48 (5 CTAs)
template <class R, class T, class O, bool BAD, int MODE>
__global__ void __launch_bounds__(RD_THREADS, rd_min_blocks<R>())
reduce_rows_kernel(const __grid_constant__ RdPlan p) {
  using Acc = typename R::Acc;
  const T abad = from_bits<T>(p.abad);
  const bool abadnan = p.abadnan != 0;
  const int chunk_id = blockIdx.y;
  const int64_t lo = (int64_t)chunk_id * p.chunk;
  const int64_t hi = (lo + p.chunk < p.n) ? lo + p.chunk : p.n;
  __shared__ Acc smem[RD_THREADS / 32 + 1];

  int64_t row, row_step; int lane, width;
  if (MODE == 0) { row = (int64_t)blockIdx.x * RD_THREADS + threadIdx.x; row_step = (int64_t)gridDim.x * RD_THREADS; lane = 0; width = 1; }
  else if (MODE == 1) { row = (int64_t)blockIdx.x * (RD_THREADS / 32) + (threadIdx.x >> 5); row_step = (int64_t)gridDim.x * (RD_THREADS / 32); lane = threadIdx.x & 31; width = 32; }
  else { row = blockIdx.x; row_step = gridDim.x; lane = threadIdx.x; width = RD_THREADS; }

  for (; row < p.nrows; row += row_step) {
    int64_t oa, ob;
    rd_row_offsets(p, row, oa, ob);
    const T *rp = reinterpret_cast<const T *>(p.a) + oa;
    typename R::Loc loc = R::linit();
    rd_row<R, T, BAD>(loc, rp, lo, hi, p.inc_n, lane, width, abad, abadnan);
    Acc mine;
    if constexpr (R::kRescan) {
      if (R::needs_rescan(loc)) {   // rare: this thread saw only identity-valued / NaN / BAD elements
        typename R::Exact::Loc ex = R::Exact::linit();
        rd_row<typename R::Exact, T, BAD>(ex, rp, lo, hi, p.inc_n, lane, width, abad, abadnan);
        mine = R::Exact::lift(ex, lo);
      } else mine = R::lift(loc, lo);
    } else mine = R::lift(loc, lo);
    Acc acc = rd_group_reduce<R, MODE>(mine, smem);
    const bool writer = (MODE == 0) || (MODE == 1 && lane == 0) || (MODE == 2 && threadIdx.x == 0);
    if (p.nchunks == 1) {
      O *out = reinterpret_cast<O *>(p.b) + ob;
      if constexpr (rd_kfind<R>::value) {   // two-phase _ind: acc holds the row's extreme VALUE, every thread has it
        static_assert(sizeof(Acc) >= sizeof(int64_t), "smem slots are reused for the index reduction");
        int64_t first = RD_NOIDX;
        if (acc.any) first = rd_first_eq<T>(rp, p.n, p.inc_n, lane, width, acc.cur);
        first = rd_group_min64<MODE>(first, reinterpret_cast<int64_t *>(smem));
        if (writer) *out = acc.any ? (O)first : from_bits<O>(p.bbad);
        continue;
      }
      if constexpr (R::kPrefix) {
        if (R::may_underflow(acc, p.n)) {   // uniform over the group: every thread holds the total
          if (rd_neg_log<T, BAD, MODE>(rp, p.n, p.inc_n, lane, width, abad, abadnan, smem) <= R::kZeroLog2) {
            if (writer) *out = rd_prod_sequential<T, O, BAD>(rp, p.n, p.inc_n, abad, abadnan, from_bits<O>(p.bbad));
            continue;
          }
        }
        if (acc.z != RD_NOIDX) {
          const O v = rd_prod_prefix<R, T, O, BAD, MODE>(rp, acc.z, p.inc_n, lane, width, abad, abadnan, smem);
          if (writer) *out = v;
          continue;
        }
      }
      if (writer) R::finish(acc, p, out);
    } else if (writer) {
      reinterpret_cast<Acc *>(p.partial)[row * p.nchunks + chunk_id] = acc;
    }
  }
}

// second stage: merges each row's partials — one warp per row, or (few rows with many partials each: the whole-array
// wrappers) one CTA per row, so that ~1200 partials are not walked by a single warp
template <class R, class T, class O, bool BAD>
__global__ void __launch_bounds__(RD_THREADS)
reduce_finish_kernel(const __grid_constant__ RdPlan p, const int cta_per_row) {
  using Acc = typename R::Acc;
  __shared__ Acc smem[RD_THREADS / 32 + 1];
  const int lane = cta_per_row ? threadIdx.x : (threadIdx.x & 31), width = cta_per_row ? RD_THREADS : 32;
  int64_t row = cta_per_row ? blockIdx.x : (int64_t)blockIdx.x * (RD_THREADS / 32) + (threadIdx.x >> 5);
  const int64_t row_step = cta_per_row ? gridDim.x : (int64_t)gridDim.x * (RD_THREADS / 32);
  for (; row < p.nrows; row += row_step) {
    int64_t oa, ob;
    rd_row_offsets(p, row, oa, ob);
    Acc acc = R::init();
    const Acc *part = reinterpret_cast<const Acc *>(p.partial) + row * p.nchunks;
    for (int c = lane; c < p.nchunks; c += width) acc = R::merge(acc, part[c]);
    if (cta_per_row) acc = rd_group_reduce<R, 2>(acc, smem);
    else {
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) acc = R::merge(acc, shfl_xor_acc(acc, m));
    }
    O *out = reinterpret_cast<O *>(p.b) + ob;
    if constexpr (R::kPrefix) {
      if (R::may_underflow(acc, p.n)) {
        const T *rp = reinterpret_cast<const T *>(p.a) + oa;
        const float negl = cta_per_row ? rd_neg_log<T, BAD, 2>(rp, p.n, p.inc_n, lane, width, from_bits<T>(p.abad), p.abadnan != 0, smem)
                                       : rd_neg_log<T, BAD, 1>(rp, p.n, p.inc_n, lane, 32, from_bits<T>(p.abad), p.abadnan != 0, nullptr);
        if (negl <= R::kZeroLog2) {
          if (lane == 0) *out = rd_prod_sequential<T, O, BAD>(rp, p.n, p.inc_n, from_bits<T>(p.abad), p.abadnan != 0, from_bits<O>(p.bbad));
          continue;
        }
      }
      if (acc.z != RD_NOIDX) {
        const T *rp = reinterpret_cast<const T *>(p.a) + oa;
        O v;
        if (cta_per_row) v = rd_prod_prefix<R, T, O, BAD, 2>(rp, acc.z, p.inc_n, lane, width, from_bits<T>(p.abad), p.abadnan != 0,
                                                             reinterpret_cast<typename RProd<T, O>::Acc *>(smem));
        else v = rd_prod_prefix<R, T, O, BAD, 1>(rp, acc.z, p.inc_n, lane, 32, from_bits<T>(p.abad), p.abadnan != 0, nullptr);
        if (lane == 0) *out = v;
        continue;
      }
    }
    if (lane == 0) R::finish(acc, p, out);
  }
}

// host planner (reduce_plan.cu)
struct RdLaunch { int mode; dim3 grid; };
int rd_build_plan(const pdlb200_trans *t, size_t in_size, size_t out_size, size_t acc_size,
                  RdPlan *p, RdLaunch *l, const Err &E, bool heavy_row_end = false, int blocks_per_sm = 8);

template <class R, class T, class O>
int rd_launch_typed(const pdlb200_trans *t, const char *name, const Err &E) {
  RdPlan p; RdLaunch l;
  int rc = rd_build_plan(t, sizeof(T), sizeof(O), sizeof(typename R::Acc), &p, &l, E, R::kRescan, rd_min_blocks<R>());
  if (rc) return rc;
  if (p.nrows == 0) return PDLB200_OK;
  cudaStream_t s = (cudaStream_t)t->stream;
#define PDLB200_RD_GO(BADF, MODE) reduce_rows_kernel<R, T, O, BADF, MODE><<<l.grid, RD_THREADS, 0, s>>>(p)
  if (t->bvalflag) { if (l.mode == 0) PDLB200_RD_GO(true, 0); else if (l.mode == 1) PDLB200_RD_GO(true, 1); else PDLB200_RD_GO(true, 2); }
  else             { if (l.mode == 0) PDLB200_RD_GO(false, 0); else if (l.mode == 1) PDLB200_RD_GO(false, 1); else PDLB200_RD_GO(false, 2); }
#undef PDLB200_RD_GO
  note_launch(name);
  PDLB200_CUDA_OK(cudaGetLastError(), E);
  if (p.nchunks > 1) {
    const int cta_per_row = (p.nrows <= 64 && p.nchunks >= 64) ? 1 : 0;
    int64_t g = cta_per_row ? p.nrows : (p.nrows + RD_THREADS / 32 - 1) / (RD_THREADS / 32);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (g > cap) g = cap;
    if (t->bvalflag) reduce_finish_kernel<R, T, O, true><<<(int)g, RD_THREADS, 0, s>>>(p, cta_per_row);
    else reduce_finish_kernel<R, T, O, false><<<(int)g, RD_THREADS, 0, s>>>(p, cta_per_row);
    note_launch("reduce_finish");
    PDLB200_CUDA_OK(cudaGetLastError(), E);
  }
  return PDLB200_OK;
}

}  // namespace pdlb200
