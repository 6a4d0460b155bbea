// elementwise.cuh — the broadcast-loop walker kernel for PDL::Ops bodies.
//
// Replaces the two innermost `for`s of PDL_BROADCASTLOOP_START plus the odometer
// of PDL_BROADCASTLOOP_END (lib/PDL/Core/pdl.h.PL:640-675,
// lib/PDL/Core/pdlbroadcast.h:72-84) for the biop / bifunc / ufunc bodies of
// lib/PDL/Ops.pd:104-265.  Roofline: HBM.  Algorithmic bytes per element:
// sizeof(T) * (distinct inputs read + 1 output written); stride-0 (dummy)
// operands cost nothing.
//
// Work unit = VEC = 16/sizeof(T) consecutive positions along collapsed dim 0
// (one 128-bit register image per operand).  A thread owns UNROLL units per trip
// of a grid-stride loop, issues all its loads first (UNROLL*NIN independent
// 128-bit LDGs in flight), then computes and stores.  Operands that are not
// unit-stride/16B-aligned along dim 0 (strided slices, reversed views, dummy
// dims) are gathered element-wise into the same register image, so vaffine views
// are read in place and never materialised.
#pragma once
#include <cstdlib>
#include <type_traits>
#include "common.cuh"

namespace pdlb200 {

struct EwPlan {
  char *ptr[3];                // operand base (offs applied), bytes; [NIN] is the output
  int64_t st[3][MAXD];         // element strides per collapsed dim
  int64_t dims[MAXD];
  int64_t n_units;             // vpr * prod(dims[1..])
  int64_t vpr;                 // vectors per row = ceil(dims[0]/VEC)   (tile kernel: tiles per row)
  int64_t ipr;                 // tile kernel: work items per row
  int grp;                     // tile kernel: consecutive tiles of one row per work item
  int rowrep;                  // tile kernel: consecutive dim-1 rows per work item (>1 only when an input is
                               //   constant along dim 1: its tile is then loaded once and reused from registers)
  int64_t nblk1;               // ceil(dims[1] / rowrep)
  uint64_t bad[3];             // badvalue bits per operand
  int nd;
  int badnan[3];
  int badchk[3];               // test this input for BAD (state flag for biop; always for bifunc/ufunc)
  int vec[3];                  // 128-bit access allowed on full units
  uint64_t param;              // bad-aware ops: bits of the OtherPars value, already cast to the element type
  int *flag;                   // bad-aware ops with a data-dependent output badflag: device word, set to 1
};

// Element types may differ per operand (convert: TI -> TO; setbadif: int mask): a unit is as many
// elements as fit 16 bytes of the WIDEST type involved.
template <class TI, class TO, class TB> constexpr int ew_vec() {
  size_t w = sizeof(TI) > sizeof(TO) ? sizeof(TI) : sizeof(TO);
  if (sizeof(TB) > w) w = sizeof(TB);
  return (int)(16 / w);
}

// Functor protocols.  Plain ops: `TO f<TI,TO>(TI a, TI b)`; the kernel handles BAD (any BAD input -> BAD out).
// Bad-aware ops (lib/PDL/Bad.pd) declare `static constexpr bool kBadAware = true` and get the BAD tests as
// arguments: `TO g<TI,TB,TO>(TI a, bool abad, TB b, bool bbad, TO cbad, uint64_t param, int &flag)`.
template <class Op, class = void> struct op_badaware : std::false_type {};
template <class Op> struct op_badaware<Op, std::void_t<decltype(Op::kBadAware)>> : std::true_type {};

template <class Op, class TI, class TB, class TO, bool BAD, int NIN>
__device__ __forceinline__ TO ew_apply(const EwPlan &p, TI a, TB b, TI abad, TB bbad, TO cbad, int &flag) {
  bool ba = false, bb = false;
  if constexpr (BAD) {
    ba = p.badchk[0] && is_bad(a, abad, p.badnan[0] != 0);
    if (NIN > 1) bb = p.badchk[1] && is_bad(b, bbad, p.badnan[1] != 0);
  }
  if constexpr (op_badaware<Op>::value) {
    return Op::template g<TI, TB, TO>(a, ba, b, bb, cbad, p.param, flag);
  } else {
    const TO r = Op::template f<TI, TO>(a, b);
    return (ba || bb) ? cbad : r;
  }
}
// ---- SIMD-within-a-register forms for 8/16-bit integer types -------------------------------------------------
// A 16-byte unit holds 16 (8) elements: per-element code costs 5-7 issue slots per element (extract, compare with
// the badvalue twice, op, select, insert), which makes these types issue-bound far below the HBM roofline in BAD
// mode.  Here the BAD test and the merge with the output badvalue are done on whole 32-bit words (the exact
// zero-lane test on `w ^ badword`), and ops that have a lane-wise word form (`Op::kPackedWords` + `fw<T>(a, b)`)
// never unpack at all.  Bit-exact: wrapping integer arithmetic lane by lane.
template <class Op, class = void> struct op_packed : std::false_type {};
template <class Op> struct op_packed<Op, std::void_t<decltype(Op::kPackedWords)>> : std::true_type {};

// float/double ops that evaluate a whole unit at once (`Op::kPackFloat` + `fpack<T, VEC>(a, b, c)`, ew_ops.cuh)
template <class Op, class = void> struct op_packfloat : std::false_type {};
template <class Op> struct op_packfloat<Op, std::void_t<decltype(Op::kPackFloat)>> : std::true_type {};

// ops whose body needs more than the 64 registers of the 4-CTAs/SM configuration (`Op::kHeavy`): 3 CTAs/SM, 80 registers
template <class Op, class = void> struct op_heavy : std::false_type {};
template <class Op> struct op_heavy<Op, std::void_t<decltype(Op::kHeavy)>> : std::true_type {};

template <class T> __device__ __forceinline__ uint32_t ew_splat(T v) {
  if constexpr (sizeof(T) == 1) return 0x01010101u * (uint32_t)(uint8_t)v; else return 0x00010001u * (uint32_t)(uint16_t)v;
}
// all-ones in every lane of w that equals the same lane of badw
template <class T> __device__ __forceinline__ uint32_t ew_eq_lanes(uint32_t w, uint32_t badw) {
  const uint32_t x = w ^ badw;
  if constexpr (sizeof(T) == 1) {
    const uint32_t hi = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;   // 0x80 in each zero byte, exact
    return (hi >> 7) * 0xffu;
  } else {
    const uint32_t hi = ~(((x & 0x7fff7fffu) + 0x7fff7fffu) | x) & 0x80008000u;
    return (hi >> 15) * 0xffffu;
  }
}

template <class Op> __device__ __forceinline__ void ew_publish_flag(const EwPlan &p, int flag) {
  if constexpr (op_badaware<Op>::value) { if (flag && p.flag) atomicOr(p.flag, 1); }
}

// Vector access moves VEC elements = VEC*sizeof(T) bytes (16 for same-type ops;
// 8/4/2/1 for the narrow side of a convert).
template <int BYTES> struct vec_word;
template <> struct vec_word<16> { using type = uint4; };
template <> struct vec_word<8>  { using type = uint2; };
template <> struct vec_word<4>  { using type = uint32_t; };
template <> struct vec_word<2>  { using type = uint16_t; };
template <> struct vec_word<1>  { using type = uint8_t; };

template <class T, int VEC>
__device__ __forceinline__ void ew_load(Pack<T> &r, const char *base, int64_t off, int64_t st0, int vec_ok, int cnt) {
  using W = typename vec_word<VEC * sizeof(T)>::type;
  const T *p = reinterpret_cast<const T *>(base) + off;
  if (vec_ok && cnt == VEC) {
    *reinterpret_cast<W *>(&r) = *reinterpret_cast<const W *>(p);
  } else if (st0 == 0) {
    T v = *p;
#pragma unroll
    for (int k = 0; k < VEC; k++) r.e[k] = v;
  } else {
#pragma unroll
    for (int k = 0; k < VEC; k++) r.e[k] = (k < cnt) ? p[k * st0] : T(0);
  }
}

template <class T, int VEC>
__device__ __forceinline__ void ew_store(const Pack<T> &r, char *base, int64_t off, int64_t st0, int vec_ok, int cnt) {
  using W = typename vec_word<VEC * sizeof(T)>::type;
  T *p = reinterpret_cast<T *>(base) + off;
  if (vec_ok && cnt == VEC) {
    *reinterpret_cast<W *>(p) = *reinterpret_cast<const W *>(&r);
  } else {
#pragma unroll
    for (int k = 0; k < VEC; k++) if (k < cnt) p[k * st0] = r.e[k];
  }
}

// Op: struct with `template<class T> static __device__ T f(T a, T b)`.
// TI = input element type, TO = output element type (differ only for convert).
template <class Op, class TI, class TO, bool BAD, int NIN, int UNROLL, class TB = TI>
__global__ void __launch_bounds__(EW_THREADS)
ew_kernel(const __grid_constant__ EwPlan p) {
  constexpr int VEC = ew_vec<TI, TO, TB>();
  int flag = 0;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const TI abad = from_bits<TI>(p.bad[0]);
  const TB bbad = from_bits<TB>(p.bad[NIN > 1 ? 1 : 0]);
  const TO cbad = from_bits<TO>(p.bad[NIN]);

  for (int64_t u0 = tid; u0 < p.n_units; u0 += nthreads * UNROLL) {
    // Pack<> is sized for 16 bytes of the WIDER type; with mixed widths (convert)
    // the narrower side simply uses the first VEC lanes of its image.
    Pack<TI> ra[UNROLL];
    Pack<TB> rb[UNROLL];
    int64_t oc[UNROLL];
    int cnt[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; j++) {
      const int64_t u = u0 + (int64_t)j * nthreads;
      cnt[j] = 0;
      if (u < p.n_units) {
        int64_t row, v;
        if (p.nd == 1) { row = 0; v = u; }
        else if ((uint64_t)p.n_units <= 0xffffffffull) {
          uint32_t r32 = (uint32_t)u / (uint32_t)p.vpr; row = r32; v = (uint32_t)u - r32 * (uint32_t)p.vpr;
        } else { row = u / p.vpr; v = u - row * p.vpr; }
        const int64_t i0 = v * VEC;
        int64_t oa = i0 * p.st[0][0], ob = (NIN > 1) ? i0 * p.st[1][0] : 0, o = i0 * p.st[NIN][0];
        for (int d = 1; d < p.nd; d++) {
          int64_t q, i;
          if (d == p.nd - 1) { i = row; q = 0; }
          else if ((uint64_t)row <= 0xffffffffull && (uint64_t)p.dims[d] <= 0xffffffffull) {
            uint32_t q32 = (uint32_t)row / (uint32_t)p.dims[d]; q = q32; i = (uint32_t)row - q32 * (uint32_t)p.dims[d];
          } else { q = row / p.dims[d]; i = row - q * p.dims[d]; }
          oa += i * p.st[0][d];
          if (NIN > 1) ob += i * p.st[1][d];
          o += i * p.st[NIN][d];
          row = q;
        }
        const int64_t left = p.dims[0] - i0;
        cnt[j] = left < VEC ? (int)left : VEC;
        oc[j] = o;
        ew_load<TI, VEC>(ra[j], p.ptr[0], oa, p.st[0][0], p.vec[0], cnt[j]);
        if (NIN > 1) ew_load<TB, VEC>(rb[j], p.ptr[1], ob, p.st[1][0], p.vec[1], cnt[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < UNROLL; j++) {
      if (cnt[j] > 0) {
        Pack<TO> rc;
#pragma unroll
        for (int k = 0; k < VEC; k++) {
          const TI a = ra[j].e[k];
          const TB b = (NIN > 1) ? rb[j].e[k] : TB(0);
          rc.e[k] = ew_apply<Op, TI, TB, TO, BAD, NIN>(p, a, b, abad, bbad, cbad, flag);
        }
        ew_store<TO, VEC>(rc, p.ptr[NIN], oc[j], p.st[NIN][0], p.vec[NIN], cnt[j]);
      }
    }
  }
  ew_publish_flag<Op>(p, flag);
}

// ---- tile kernel: the fast path --------------------------------------------------------------
// A tile = TILE = EW_THREADS*UNROLL*VEC consecutive positions of collapsed dim 0 inside ONE outer
// row; a work item = up to `grp` consecutive tiles of one row.  The outer-index decode (div/mod
// over dims[1..]) and the per-operand row offsets are computed once per work item; inside it every
// unit is base + constant stride.  Stride-0 (dummy-dim) operands are read ONCE per work item and
// replicated in registers; tiles that lie completely inside the row take a body without bounds
// checks (FULL), only the last tile of a row pays for them.
// Used when dim 0 is long enough to fill tiles (always for fully collapsed contiguous ndarrays).
template <class T, int VEC>
__device__ __forceinline__ void ew_load_full(Pack<T> &r, const T *p, int64_t st0, int vec_ok, T bc, bool is_bc) {
  using W = typename vec_word<VEC * sizeof(T)>::type;
  if (is_bc) {
#pragma unroll
    for (int k = 0; k < VEC; k++) r.e[k] = bc;
    return;
  }
  if (vec_ok) { *reinterpret_cast<W *>(&r) = *reinterpret_cast<const W *>(p); return; }
#pragma unroll
  for (int k = 0; k < VEC; k++) r.e[k] = p[k * st0];
}

template <class Op, class TI, class TO, bool BAD, int NIN, int VEC, class TB>
__device__ __forceinline__ void ew_compute_store(const EwPlan &p, const Pack<TI> &ra, const Pack<TB> &rb, TO *dst, int64_t sc0,
                                                 int cnt, TI abad, TB bbad, TO cbad, int &flag) {
  Pack<TO> rc;
  constexpr bool kSwar = tt<TI>::is_int && sizeof(TI) <= 2 && std::is_same<TI, TO>::value && std::is_same<TI, TB>::value &&
                         !op_badaware<Op>::value && (BAD || op_packed<Op>::value);
  if constexpr (kSwar) {
    const uint32_t aw[4] = {ra.q.x, ra.q.y, ra.q.z, ra.q.w};
    const uint32_t bw[4] = {rb.q.x, rb.q.y, rb.q.z, rb.q.w};
    uint32_t cw[4];
    if constexpr (op_packed<Op>::value) {
#pragma unroll
      for (int i = 0; i < 4; i++) cw[i] = Op::template fw<TI>(aw[i], (NIN > 1) ? bw[i] : 0u);
    } else {
#pragma unroll
      for (int k = 0; k < VEC; k++) rc.e[k] = Op::template f<TI, TO>(ra.e[k], (NIN > 1) ? rb.e[k] : TB(0));
      cw[0] = rc.q.x; cw[1] = rc.q.y; cw[2] = rc.q.z; cw[3] = rc.q.w;
    }
    if constexpr (BAD) {
      const uint32_t abw = ew_splat<TI>(abad), bbw = ew_splat<TB>(bbad), cbw = ew_splat<TO>(cbad);
      const uint32_t ca = p.badchk[0] ? 0xffffffffu : 0u, cb = (NIN > 1 && p.badchk[1]) ? 0xffffffffu : 0u;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        uint32_t m = ew_eq_lanes<TI>(aw[i], abw) & ca;
        if (NIN > 1) m |= ew_eq_lanes<TB>(bw[i], bbw) & cb;
        cw[i] = (cw[i] & ~m) | (cbw & m);
      }
    }
    rc.q = make_uint4(cw[0], cw[1], cw[2], cw[3]);
  } else if constexpr (op_packfloat<Op>::value && !tt<TI>::is_int && std::is_same<TI, TO>::value && std::is_same<TI, TB>::value) {
    if constexpr (BAD) {
      // BAD lanes (type-min / huge badvalues) would push the whole unit onto the op's rare path: give them a
      // harmless operand pair (1 op 1) first, then put the output badvalue in their place
      Pack<TI> ma; Pack<TB> mb;
      bool bad[VEC];
      // one compare per operand element whatever the mode: against the badvalue, against the element itself when the
      // badvalue is NaN (bad <=> the element is not equal to itself), against NaN when the operand is not checked
      const bool nan_a = p.badchk[0] && p.badnan[0] != 0, nan_b = NIN > 1 && p.badchk[1] && p.badnan[1] != 0;
      const TI cmp_a = p.badchk[0] ? abad : TI(NAN);
      const TB cmp_b = (NIN > 1 && p.badchk[1]) ? bbad : TB(NAN);
#pragma unroll
      for (int k = 0; k < VEC; k++) {
        bad[k] = (ra.e[k] == (nan_a ? ra.e[k] : cmp_a)) != nan_a;
        if (NIN > 1) bad[k] = bad[k] || ((rb.e[k] == (nan_b ? rb.e[k] : cmp_b)) != nan_b);
        ma.e[k] = bad[k] ? TI(1) : ra.e[k];
        mb.e[k] = (NIN > 1) ? (bad[k] ? TB(1) : rb.e[k]) : TB(1);
      }
      Op::template fpack<TI, VEC>(ma, mb, rc);
#pragma unroll
      for (int k = 0; k < VEC; k++) rc.e[k] = bad[k] ? cbad : rc.e[k];
    } else {
      Op::template fpack<TI, VEC>(ra, rb, rc);
    }
  } else {
#pragma unroll
    for (int k = 0; k < VEC; k++) {
      const TI a = ra.e[k];
      const TB b = (NIN > 1) ? rb.e[k] : TB(0);
      rc.e[k] = ew_apply<Op, TI, TB, TO, BAD, NIN>(p, a, b, abad, bbad, cbad, flag);
    }
  }
  ew_store<TO, VEC>(rc, reinterpret_cast<char *>(dst), 0, sc0, p.vec[NIN], cnt);
}

// 64 registers (4 CTAs/SM) for the good-mode bodies; the BAD bodies carry the extra compares/selects and
// get 80 (3 CTAs/SM) rather than spilling inside the hot loop.
template <class Op, class TI, class TO, bool BAD, int NIN, int UNROLL, class TB = TI>
__global__ void __launch_bounds__(EW_THREADS, (BAD || op_heavy<Op>::value) ? 3 : 4)
ew_tile_kernel(const __grid_constant__ EwPlan p) {
  constexpr int VEC = ew_vec<TI, TO, TB>();
  constexpr int64_t TILE = (int64_t)EW_THREADS * UNROLL * VEC;
  int flag = 0;
  const TI abad = from_bits<TI>(p.bad[0]);
  const TB bbad = from_bits<TB>(p.bad[NIN > 1 ? 1 : 0]);
  const TO cbad = from_bits<TO>(p.bad[NIN]);
  const int64_t tpr = p.vpr, ipr = p.ipr, nitems = p.n_units;
  const int64_t sa0 = p.st[0][0], sb0 = (NIN > 1) ? p.st[1][0] : 0, sc0 = p.st[NIN][0];
  const int64_t ja = (int64_t)EW_THREADS * VEC * sa0, jb = (int64_t)EW_THREADS * VEC * sb0, jc = (int64_t)EW_THREADS * VEC * sc0;
  const bool a_bc = (sa0 == 0) && !p.vec[0], b_bc = (NIN > 1) && (sb0 == 0) && !p.vec[1];

  const int64_t sa1 = (p.nd > 1) ? p.st[0][1] : 0, sb1 = (NIN > 1 && p.nd > 1) ? p.st[1][1] : 0, sc1 = (p.nd > 1) ? p.st[NIN][1] : 0;
  const bool a_inv = p.rowrep > 1 && sa1 == 0, b_inv = (NIN > 1) && p.rowrep > 1 && sb1 == 0;

  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    int64_t row = 0, it = item;
    int64_t oa = 0, ob = 0, oc = 0;
    int rows_here = 1;
    if (p.nd > 1) {
      row = item / ipr; it = item - row * ipr;
      int64_t r = row;
      {   // dim 1 is walked in blocks of `rowrep` rows
        const int64_t q = (p.nd == 2) ? 0 : r / p.nblk1;
        const int64_t i1 = (r - q * p.nblk1) * p.rowrep;
        const int64_t rem = p.dims[1] - i1;
        rows_here = rem < p.rowrep ? (int)rem : p.rowrep;
        oa += i1 * sa1; if (NIN > 1) ob += i1 * sb1; oc += i1 * sc1;
        r = q;
      }
      for (int d = 2; d < p.nd; d++) {
        const int64_t q = (d == p.nd - 1) ? 0 : r / p.dims[d];
        const int64_t i = r - q * p.dims[d];
        oa += i * p.st[0][d];
        if (NIN > 1) ob += i * p.st[1][d];
        oc += i * p.st[NIN][d];
        r = q;
      }
    }
    const int64_t seg0 = it * p.grp;
    const int64_t seg1 = (seg0 + p.grp < tpr) ? seg0 + p.grp : tpr;
    int64_t i_first = seg0 * TILE + (int64_t)threadIdx.x * VEC;   // this thread's first position in the row
    const TI *pa = reinterpret_cast<const TI *>(p.ptr[0]) + oa + i_first * sa0;
    const TB *pb = reinterpret_cast<const TB *>(p.ptr[NIN > 1 ? 1 : 0]) + ob + i_first * sb0;
    TO *pc = reinterpret_cast<TO *>(p.ptr[NIN]) + oc + i_first * sc0;

    for (int64_t seg = seg0; seg < seg1; seg++, i_first += TILE, pa += UNROLL * ja, pb += UNROLL * jb, pc += UNROLL * jc) {
      if ((seg + 1) * TILE <= p.dims[0]) {
        // full tile: no bounds checks; all loads of the tile are issued before the first use.  Inputs that
        // are constant along dim 1 are loaded for the first row of the block only.
        Pack<TI> ra[UNROLL];
        Pack<TB> rb[UNROLL];
        for (int rr = 0; rr < rows_here; rr++) {
          const TI *qa = pa + rr * sa1;
          const TB *qb = pb + rr * sb1;
          if (rr == 0 || !a_inv) {
            const TI bca = a_bc ? *qa : TI(0);    // dummy along dim 0: one load per row
#pragma unroll
            for (int j = 0; j < UNROLL; j++) ew_load_full<TI, VEC>(ra[j], qa + j * ja, sa0, p.vec[0], bca, a_bc);
          }
          if (NIN > 1 && (rr == 0 || !b_inv)) {
            const TB bcb = b_bc ? *qb : TB(0);
#pragma unroll
            for (int j = 0; j < UNROLL; j++) ew_load_full<TB, VEC>(rb[j], qb + j * jb, sb0, p.vec[1], bcb, b_bc);
          }
#pragma unroll
          for (int j = 0; j < UNROLL; j++)
            ew_compute_store<Op, TI, TO, BAD, NIN, VEC, TB>(p, ra[j], rb[NIN > 1 ? j : 0], pc + rr * sc1 + j * jc, sc0, VEC, abad, bbad, cbad, flag);
        }
      } else {
        // last (partial) tile of a row: one unit at a time
        for (int rr = 0; rr < rows_here; rr++) {
#pragma unroll 1
          for (int j = 0; j < UNROLL; j++) {
            const int64_t left = p.dims[0] - i_first - (int64_t)j * EW_THREADS * VEC;
            if (left <= 0) break;
            const int cnt = left < VEC ? (int)left : VEC;
            Pack<TI> ra;
            Pack<TB> rb;
            ew_load<TI, VEC>(ra, reinterpret_cast<const char *>(pa + rr * sa1 + j * ja), 0, sa0, p.vec[0], cnt);
            if (NIN > 1) ew_load<TB, VEC>(rb, reinterpret_cast<const char *>(pb + rr * sb1 + j * jb), 0, sb0, p.vec[1], cnt);
            ew_compute_store<Op, TI, TO, BAD, NIN, VEC, TB>(p, ra, rb, pc + rr * sc1 + j * jc, sc0, cnt, abad, bbad, cbad, flag);
          }
        }
      }
    }
  }
  ew_publish_flag<Op>(p, flag);
}

// host side: build the plan and launch (ew_plan.cu)
int ew_build_plan(const pdlb200_trans *t, int nin, size_t in_size, size_t out_size,
                  bool state_checked_bad, EwPlan *p, const Err &E, size_t b_size = 0);
int ew_grid(int64_t n_units, int unroll, const void *kernel);

template <class Op, class TI, class TO, int NIN, class TB = TI>
int ew_launch_typed(const pdlb200_trans *t, bool state_checked_bad, const char *name, const Err &E,
                    uint64_t param = 0, int *flag = nullptr) {
  EwPlan p;
  int rc = ew_build_plan(t, NIN, sizeof(TI), sizeof(TO), state_checked_bad, &p, E, sizeof(TB));
  if (rc) return rc;
  if (p.n_units == 0) return PDLB200_OK;  // empty broadcast: the loop body never runs
  p.param = param; p.flag = flag;
  cudaStream_t s = (cudaStream_t)t->stream;
  constexpr int VEC = ew_vec<TI, TO, TB>();
  constexpr int TU = 4;                                    // units per thread per tile
  constexpr int64_t TILE = (int64_t)EW_THREADS * TU * VEC;
  // Small contiguous problems (fewer than ~16 tiles per SM, e.g. cfg1's 2048x2048 doubles = 14 per SM) run the
  // grid-stride kernel on exactly one resident wave of CTAs: every thread gets the same number of units (+-1), so
  // there is no partial last wave of tiles.  Measured on cfg1: 17.3 us against 17.9 us (a 4-unit variant of the
  // grid-stride kernel was no faster).  PDLB200_EW_FLAT=<tiles per SM> moves the threshold, 0 disables it.
  static const int flat_mode = getenv("PDLB200_EW_FLAT") ? atoi(getenv("PDLB200_EW_FLAT")) : 16;
  const bool small_flat = flat_mode > 0 && p.nd == 1 && p.dims[0] < (int64_t)flat_mode * TILE * sm_count();
  if (!small_flat && (p.nd == 1 || p.dims[0] >= TILE / 2)) {
    // tile kernel: re-purpose vpr/n_units as tiles-per-row / number of work items.  A work item is
    // `grp` consecutive tiles of one row: 1 for small problems (keeps every SM busy), up to 8 for
    // big ones (amortises the per-item index decode).
    const int64_t tpr = (p.dims[0] + TILE - 1) / TILE;
    const int64_t rows = p.n_units / p.vpr;
    int64_t grp = (tpr * rows) / ((int64_t)sm_count() * 16);
    if (grp > 8) grp = 8;
    if (grp > tpr) grp = tpr;
    if (grp < 1) grp = 1;
    // an input that is constant along dim 1 (dummy dim / row vector) is kept in registers across a block of rows
    p.rowrep = 1; p.nblk1 = (p.nd > 1) ? p.dims[1] : 1;
    int64_t rowblocks = rows;
    if (p.nd > 1 && p.dims[1] > 1 && (p.st[0][1] == 0 || (NIN > 1 && p.st[1][1] == 0))) {
      int64_t rep = p.dims[1] < 16 ? p.dims[1] : 16;
      // keep enough work items to fill the GPU
      while (rep > 1 && tpr * (rows / p.dims[1]) * ((p.dims[1] + rep - 1) / rep) < (int64_t)sm_count() * 8) rep /= 2;
      if (rep > 1) {
        p.rowrep = (int)rep; p.nblk1 = (p.dims[1] + rep - 1) / rep;
        rowblocks = (rows / p.dims[1]) * p.nblk1;
        grp = 1;
      }
    }
    p.grp = (int)grp;
    p.ipr = (tpr + grp - 1) / grp;
    p.vpr = tpr; p.n_units = p.ipr * rowblocks;
    const int64_t cap = (int64_t)sm_count() * 32;
    const int grid = (int)(p.n_units < cap ? p.n_units : cap);
    if (t->bvalflag) ew_tile_kernel<Op, TI, TO, true, NIN, TU, TB><<<grid, EW_THREADS, 0, s>>>(p);
    else ew_tile_kernel<Op, TI, TO, false, NIN, TU, TB><<<grid, EW_THREADS, 0, s>>>(p);
  } else {
    constexpr int UNROLL = 2;
    if (t->bvalflag) {
      auto k = ew_kernel<Op, TI, TO, true, NIN, UNROLL, TB>;
      k<<<ew_grid(p.n_units, UNROLL, (const void *)k), EW_THREADS, 0, s>>>(p);
    } else {
      auto k = ew_kernel<Op, TI, TO, false, NIN, UNROLL, TB>;
      k<<<ew_grid(p.n_units, UNROLL, (const void *)k), EW_THREADS, 0, s>>>(p);
    }
  }
  note_launch(name);
  PDLB200_CUDA_OK(cudaGetLastError(), E);
  return PDLB200_OK;
}

// dispatch a same-type op over the 11 device types; IND shares int64_t with LL.
#define PDLB200_EW_CASES_INT(OP, NIN, SC, NAME) \
  case PDLB200_SB:  return ew_launch_typed<OP, int8_t,   int8_t,   NIN>(t, SC, NAME, E); \
  case PDLB200_B:   return ew_launch_typed<OP, uint8_t,  uint8_t,  NIN>(t, SC, NAME, E); \
  case PDLB200_S:   return ew_launch_typed<OP, int16_t,  int16_t,  NIN>(t, SC, NAME, E); \
  case PDLB200_US:  return ew_launch_typed<OP, uint16_t, uint16_t, NIN>(t, SC, NAME, E); \
  case PDLB200_L:   return ew_launch_typed<OP, int32_t,  int32_t,  NIN>(t, SC, NAME, E); \
  case PDLB200_UL:  return ew_launch_typed<OP, uint32_t, uint32_t, NIN>(t, SC, NAME, E); \
  case PDLB200_IND: case PDLB200_LL: return ew_launch_typed<OP, int64_t, int64_t, NIN>(t, SC, NAME, E); \
  case PDLB200_ULL: return ew_launch_typed<OP, uint64_t, uint64_t, NIN>(t, SC, NAME, E);
#define PDLB200_EW_CASES_FLT(OP, NIN, SC, NAME) \
  case PDLB200_F:   return ew_launch_typed<OP, float,  float,  NIN>(t, SC, NAME, E); \
  case PDLB200_D:   return ew_launch_typed<OP, double, double, NIN>(t, SC, NAME, E);

}  // namespace pdlb200
