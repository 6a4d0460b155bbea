// matmult_tma.cu — double-precision matmult on the FP64 tensor cores with TMA-staged operand tiles.
//
// lib/PDL/Primitive.pd:191-264 for double, no BAD values, t unit-stride in a and w unit-stride in b (PDL's
// physical layout), one matrix pair per call.  Same tiling and DMMA inner loop as matmult_dmma.cu (CTA tile
// 128x128, BK = 16, 16 consumer warps of 32x32), but the operand tiles are moved by the TMA engine: ONE elected
// thread (lane 0 of consumer warp 0, between two of its own k-tiles) issues `cp.async.bulk.tensor.2d` (SASS
// UTMALDG) per box and arms the stage's mbarrier with `mbarrier.arrive.expect_tx`; the four LDGSTS producer warps
// of the cp.async version (2048 copies per stage, their address arithmetic, their issue slots and their registers)
// are gone — 16 warps, 128 registers each — and out-of-range parts of edge tiles are zero-filled by the hardware
// instead of by predicated copies.
//
// Shared-memory layout = what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B: rows of 128 bytes (16 doubles), the
// 16-byte chunk index XORed with (row & 7).  A tile: one box {16 k, 128 h}.  B tile: eight boxes {16 w, 16 k}.
// The 64-bit fragment loads are bank-conflict free WITHOUT padding because the kernel picks, inside each m8n8k4
// step, WHICH rows and WHICH k the lanes hold (the MMA only needs A and B to agree on k, and C to know its rows):
//   lane g of an m8 tile holds row   PERM[g] = {0,1,4,5,2,3,6,7}      -> the two half-warps hit disjoint chunk pairs
//   lane t of k-step s     holds k = 8*(s>>1) + {0,3,4,7} or {1,2,5,6} -> four rows of B with distinct (k & 7) >> 1
// (derivation in DESIGN.md §4.3).  Roofline: FP64 tensor (DMMA) peak; 2*T*H*W flop.
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include "matmult.cuh"

namespace pdlb200 {

constexpr int TM_BM = 128, TM_BN = 128, TM_BK = 16;
constexpr int TM_STAGES = 5;
constexpr int TM_A_BYTES = TM_BM * TM_BK * 8;          // 16 KB
constexpr int TM_B_BYTES = TM_BK * TM_BN * 8;          // 16 KB = 8 boxes of 2 KB
constexpr int TM_STAGE_BYTES = TM_A_BYTES + TM_B_BYTES;
constexpr int TM_NCW = 16;                             // consumer warps
constexpr size_t TM_SMEM = (size_t)TM_STAGES * TM_STAGE_BYTES + 1024 /* alignment slack */ + 2 * TM_STAGES * sizeof(uint64_t);
constexpr int TM_THREADS = TM_NCW * 32;                // 512: no dedicated producer warp, 128 registers per thread

__device__ __forceinline__ void tm_mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void tm_mbar_arrive(uint64_t *bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" :: "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void tm_mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n"
               :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tm_mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n"
      :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
// one box of a 2-d tensor map -> shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void tm_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
               :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)),
                  "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tm_dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void tm_issue_stage(unsigned char *smem, uint64_t *full, const CUtensorMap *mapA, const CUtensorMap *mapB,
                                               int kt, int h0, int w0) {
  const int s = kt % TM_STAGES;
  unsigned char *dA = smem + (size_t)s * TM_STAGE_BYTES, *dB = dA + TM_A_BYTES;
  tm_mbar_expect_tx(&full[s], TM_STAGE_BYTES);
  tm_load_2d(dA, mapA, kt * TM_BK, h0, &full[s]);                                                   // box {16 k, 128 h}
#pragma unroll
  for (int b = 0; b < 8; b++) tm_load_2d(dB + b * 2048, mapB, w0 + 16 * b, kt * TM_BK, &full[s]);   // boxes {16 w, 16 k}
}

__global__ void __launch_bounds__(TM_THREADS, 1)
mm_dmma_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ MmPlan p) {
  extern __shared__ unsigned char tm_smem_raw[];
  // SWIZZLE_128B: the XOR uses address bits 7..9, so every tile starts on a 1024-byte boundary
  // (an offset into the __shared__ array, so that the compiler keeps the shared address space: LDS, not generic LD)
  unsigned char *smem = tm_smem_raw + ((1024u - ((unsigned)__cvta_generic_to_shared(tm_smem_raw) & 1023u)) & 1023u);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)TM_STAGES * TM_STAGE_BYTES);
  uint64_t *empty = full + TM_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h0 = blockIdx.y * TM_BM, w0 = blockIdx.x * TM_BN;
  const int KT = (int)((p.T + TM_BK - 1) / TM_BK);

  if (tid == 0) {
    for (int s = 0; s < TM_STAGES; s++) { tm_mbar_init(&full[s], 1); tm_mbar_init(&empty[s], TM_NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("prefetch.tensormap [%0];\n" :: "l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" :: "l"(&mapB) : "memory");
  }
  __syncthreads();
  // prologue: the first STAGES k-tiles are in flight before anyone computes
  if (tid == 0)
    for (int kt = 0; kt < TM_STAGES && kt < KT; kt++) tm_issue_stage(smem, full, &mapA, &mapB, kt, h0, w0);

  // ===== consumers: 4 x 4 warps, warp tile 32 x 32 =====
  const int wm = warp >> 2, wn = warp & 3;
  const int g = lane >> 2, t4 = lane & 3;
  const int prow = ((g & 2) << 1) | (g & 1) | ((g & 4) >> 1);      // PERM[g]: the row of an m8 tile this lane holds
  // per-lane byte offsets of the four k-steps of a k-tile
  unsigned aoff[4], boff[4][2];
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int kk = 8 * (s >> 1) + ((s & 1) ? ((t4 == 0) ? 1 : (t4 == 1) ? 2 : (t4 == 2) ? 5 : 6)
                                           : ((t4 == 0) ? 0 : (t4 == 1) ? 3 : (t4 == 2) ? 4 : 7));
    aoff[s] = (unsigned)(prow * 128 + ((((kk >> 1) ^ prow) & 7) << 4) + ((kk & 1) << 3));
#pragma unroll
    for (int jj = 0; jj < 2; jj++)
      boff[s][jj] = (unsigned)(kk * 128 + ((((jj * 4 + (g >> 1)) ^ kk) & 7) << 4) + ((g & 1) << 3));
  }
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

  for (int kt = 0; kt < KT; kt++) {
    const int s = kt % TM_STAGES;
    if (tid == 0 && kt >= 1) {
      // refill the slot k-tile kt-1 used, once all 16 warps have released it (this warp did at the end of kt-1)
      const int nk = kt - 1 + TM_STAGES;
      if (nk < KT) {
        tm_mbar_wait(&empty[nk % TM_STAGES], ((nk / TM_STAGES) - 1) & 1);
        tm_issue_stage(smem, full, &mapA, &mapB, nk, h0, w0);
      }
    }
    __syncwarp();
    tm_mbar_wait(&full[s], (kt / TM_STAGES) & 1);
    const unsigned char *tA = smem + (size_t)s * TM_STAGE_BYTES + (wm * 32) * 128;
    const unsigned char *tB = smem + (size_t)s * TM_STAGE_BYTES + TM_A_BYTES + (wn * 2) * 2048;
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; i++) af[i] = *reinterpret_cast<const double *>(tA + i * 1024 + aoff[ks]);
#pragma unroll
      for (int j = 0; j < 4; j++) bf[j] = *reinterpret_cast<const double *>(tB + (j >> 1) * 2048 + boff[ks][j & 1]);
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) tm_dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncwarp();
    if (lane == 0) tm_mbar_arrive(&empty[s]);
  }

  double *C = reinterpret_cast<double *>(p.c);
  const bool c_vec = (p.icw == 1) && ((((uintptr_t)C) & 15) == 0) && ((p.ich & 1) == 0);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int64_t h = (int64_t)h0 + wm * 32 + i * 8 + prow;
    if (h >= p.H) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int64_t w = (int64_t)w0 + wn * 32 + j * 8 + t4 * 2;
      if (w >= p.W) continue;
      double *dst = C + h * p.ich + w * p.icw;
      if (c_vec && w + 1 < p.W) *reinterpret_cast<double2 *>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
      else { dst[0] = acc[i][j][0]; if (w + 1 < p.W) dst[p.icw] = acc[i][j][1]; }
    }
  }
}

// ===== stream-K version: persistent CTAs, the (tile, k-tile) iteration space cut evenly across them =====
// A grid of 128x128 tiles only fills the machine in whole waves: 2048^3 is 256 tiles on 148 SMs = 1.73 waves that take
// the time of 2.  Here G <= #SMs persistent CTAs each take a CONTIGUOUS range of the global iteration space
// (iteration = one k-tile of one output tile, tiles in row-major order), so every SM does the same amount of DMMA
// work and the TMA ring never drains between tiles.  A tile whose k range is cut is finished by the CTA that holds its
// k = 0 end: that CTA reaches the tile LAST in its own range, whereas the CTAs holding the rest of the tile reach it
// FIRST — they have written their partial accumulators (thread-linear layout, 128 KB per CTA, L2-resident) and
// raised their flag long before the finisher asks; all CTAs are resident (one per SM), so the wait cannot deadlock.
// The partial sums are added in k order: the result does not depend on timing.
struct SkPlan {
  double *ws;          // [G][32][512] partial accumulators
  int *flags;          // [G], zeroed per launch
  int KT, ntx, nty, G;
  long long total;     // ntx * nty * KT iterations
};

__device__ __forceinline__ void sk_range(const SkPlan &k, int b, long long &lo, long long &hi) {
  lo = k.total * b / k.G; hi = k.total * (b + 1) / k.G;
}

__global__ void __launch_bounds__(TM_THREADS, 1)
mm_dmma_tma_sk_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ MmPlan p,
                      const __grid_constant__ SkPlan k) {
  extern __shared__ unsigned char tm_smem_raw[];
  unsigned char *smem = tm_smem_raw + ((1024u - ((unsigned)__cvta_generic_to_shared(tm_smem_raw) & 1023u)) & 1023u);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)TM_STAGES * TM_STAGE_BYTES);
  uint64_t *empty = full + TM_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  long long lo, hi;
  sk_range(k, blockIdx.x, lo, hi);
  const int nloc = (int)(hi - lo);                       // iterations of this CTA

  if (tid == 0) {
    for (int s = 0; s < TM_STAGES; s++) { tm_mbar_init(&full[s], 1); tm_mbar_init(&empty[s], TM_NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("prefetch.tensormap [%0];\n" :: "l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" :: "l"(&mapB) : "memory");
  }
  __syncthreads();
  // iterations are issued in order: the coordinates of the NEXT one are kept incrementally (no divisions on the
  // issuing lane's path — it is also a consumer, and the slowest warp paces the ring)
  int p_k, p_h, p_w;
  {
    const int tile = (int)(lo / k.KT);
    p_k = (int)(lo - (long long)tile * k.KT) * TM_BK;
    p_h = (tile / k.ntx) * TM_BM; p_w = (tile % k.ntx) * TM_BN;
  }
  const int k_end = k.KT * TM_BK, w_end = k.ntx * TM_BN;
  auto issue = [&](int q) {
    const int s = q % TM_STAGES;
    unsigned char *dA = smem + (size_t)s * TM_STAGE_BYTES, *dB = dA + TM_A_BYTES;
    tm_mbar_expect_tx(&full[s], TM_STAGE_BYTES);
    tm_load_2d(dA, &mapA, p_k, p_h, &full[s]);
#pragma unroll
    for (int b = 0; b < 8; b++) tm_load_2d(dB + b * 2048, &mapB, p_w + 16 * b, p_k, &full[s]);
    p_k += TM_BK;
    if (p_k >= k_end) { p_k = 0; p_w += TM_BN; if (p_w >= w_end) { p_w = 0; p_h += TM_BM; } }
  };
  if (tid == 0)
    for (int q = 0; q < TM_STAGES && q < nloc; q++) issue(q);

  const int wm = warp >> 2, wn = warp & 3;
  const int g = lane >> 2, t4 = lane & 3;
  const int prow = ((g & 2) << 1) | (g & 1) | ((g & 4) >> 1);
  unsigned aoff[4], boff[4][2];
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int kk = 8 * (s >> 1) + ((s & 1) ? ((t4 == 0) ? 1 : (t4 == 1) ? 2 : (t4 == 2) ? 5 : 6)
                                           : ((t4 == 0) ? 0 : (t4 == 1) ? 3 : (t4 == 2) ? 4 : 7));
    aoff[s] = (unsigned)(prow * 128 + ((((kk >> 1) ^ prow) & 7) << 4) + ((kk & 1) << 3));
#pragma unroll
    for (int jj = 0; jj < 2; jj++)
      boff[s][jj] = (unsigned)(kk * 128 + ((((jj * 4 + (g >> 1)) ^ kk) & 7) << 4) + ((g & 1) << 3));
  }
  double *C = reinterpret_cast<double *>(p.c);
  const bool c_vec = (p.icw == 1) && ((((uintptr_t)C) & 15) == 0) && ((p.ich & 1) == 0);

  int q = 0;
  while (q < nloc) {
    const long long g0 = lo + q;
    const int tile = (int)(g0 / k.KT), kt0 = (int)(g0 - (long long)tile * k.KT);
    const int nk = (k.KT - kt0 < nloc - q) ? k.KT - kt0 : nloc - q;    // k-tiles of this segment
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    for (int it = 0; it < nk; it++, q++) {
      const int s = q % TM_STAGES;
      if (tid == 0 && q >= 1) {
        const int nq = q - 1 + TM_STAGES;                // refill the slot iteration q-1 used
        if (nq < nloc) {
          tm_mbar_wait(&empty[nq % TM_STAGES], ((nq / TM_STAGES) - 1) & 1);
          issue(nq);
        }
      }
      __syncwarp();
      tm_mbar_wait(&full[s], (q / TM_STAGES) & 1);
      const unsigned char *tA = smem + (size_t)s * TM_STAGE_BYTES + (wm * 32) * 128;
      const unsigned char *tB = smem + (size_t)s * TM_STAGE_BYTES + TM_A_BYTES + (wn * 2) * 2048;
#pragma unroll
      for (int ks = 0; ks < 4; ks++) {
        double af[4], bf[4];
#pragma unroll
        for (int i = 0; i < 4; i++) af[i] = *reinterpret_cast<const double *>(tA + i * 1024 + aoff[ks]);
#pragma unroll
        for (int j = 0; j < 4; j++) bf[j] = *reinterpret_cast<const double *>(tB + (j >> 1) * 2048 + boff[ks][j & 1]);
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) tm_dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
      __syncwarp();
      if (lane == 0) tm_mbar_arrive(&empty[s]);
    }

    if (kt0 != 0) {
      // the later part of a tile that another CTA finishes: publish the partial accumulators
      double *w = k.ws + (size_t)blockIdx.x * (32 * TM_THREADS) + tid;
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { w[((i * 4 + j) * 2) * TM_THREADS] = acc[i][j][0]; w[((i * 4 + j) * 2 + 1) * TM_THREADS] = acc[i][j][1]; }
      __threadfence();
      __syncthreads();
      if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;\n" :: "l"(k.flags + blockIdx.x), "r"(1) : "memory");
      continue;
    }
    if (kt0 + nk < k.KT) {
      // this CTA holds the k = 0 end of a cut tile: add the parts of the CTAs after it, in k order
      int done = nk;
      for (int peer = blockIdx.x + 1; done < k.KT; peer++) {
        long long plo, phi;
        sk_range(k, peer, plo, phi);
        const long long tile_end = (long long)(tile + 1) * k.KT;
        done += (int)((phi < tile_end ? phi : tile_end) - plo);
        if (tid == 0) {
          int v;
          do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(k.flags + peer) : "memory"); } while (v == 0);
        }
        __syncthreads();
        const double *w = k.ws + (size_t)peer * (32 * TM_THREADS) + tid;
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) {
            acc[i][j][0] += __ldcg(w + ((i * 4 + j) * 2) * TM_THREADS);
            acc[i][j][1] += __ldcg(w + ((i * 4 + j) * 2 + 1) * TM_THREADS);
          }
      }
    }
    const int by = tile / k.ntx, bx = tile - by * k.ntx;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int64_t h = (int64_t)by * TM_BM + wm * 32 + i * 8 + prow;
      if (h >= p.H) continue;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int64_t w = (int64_t)bx * TM_BN + wn * 32 + j * 8 + t4 * 2;
        if (w >= p.W) continue;
        double *dst = C + h * p.ich + w * p.icw;
        if (c_vec && w + 1 < p.W) *reinterpret_cast<double2 *>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
        else { dst[0] = acc[i][j][0]; if (w + 1 < p.W) dst[p.icw] = acc[i][j][1]; }
      }
    }
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time dependency on libcuda
typedef CUresult (*tm_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tm_encode_fn tm_encoder() {
  static tm_encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (tm_encode_fn)sym;
    else cudaGetLastError();
  }
  return fn;
}

// rows of `cols` doubles, `rows` of them, `pitch` doubles apart; box {bc, br}
static bool tm_make_map(CUtensorMap *m, const void *base, int64_t cols, int64_t rows, int64_t pitch, int bc, int br) {
  tm_encode_fn enc = tm_encoder();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch * 8};
  const cuuint32_t box[2] = {(cuuint32_t)bc, (cuuint32_t)br};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// returns PDLB200_EUNSUPPORTED when the call is not eligible (the cp.async kernel takes it)
int launch_matmult_tma(const pdlb200_trans *t, const MmPlan &p, const Err &E) {
  if (const char *v = getenv("PDLB200_DMMA")) { if (strcmp(v, "tma")) return PDLB200_EUNSUPPORTED; }
  if (p.nbatch != 1 || p.T == 0 || p.iat != 1 || p.ibw != 1) return PDLB200_EUNSUPPORTED;
  if (p.H * p.W < 128 * 128) return PDLB200_EUNSUPPORTED;
  if (p.T > 0x7fffffff || p.H > 0x7fffffff || p.W > 0x7fffffff) return PDLB200_EUNSUPPORTED;
  // TMA: 16-byte aligned base and row pitch; a row pitch of 0 (size-1 dim) is not a tensor
  if ((((uintptr_t)p.a | (uintptr_t)p.b) & 15) || (p.iah & 1) || (p.ibt & 1) || p.iah < p.T || p.ibt < p.W) return PDLB200_EUNSUPPORTED;
  CUtensorMap mapA, mapB;
  if (!tm_make_map(&mapA, p.a, p.T, p.H, p.iah, TM_BK, TM_BM) || !tm_make_map(&mapB, p.b, p.W, p.T, p.ibt, 16, TM_BK))
    return PDLB200_EUNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    PDLB200_CUDA_OK(cudaFuncSetAttribute(mm_dmma_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TM_SMEM), E);
    attr_set = true;
  }
  dim3 grid((unsigned)((p.W + TM_BN - 1) / TM_BN), (unsigned)((p.H + TM_BM - 1) / TM_BM), 1);
  cudaStream_t st = (cudaStream_t)t->stream;
  // stream-K unless the tile grid already fills whole waves (or PDLB200_MM_STREAMK=0)
  static const int sk_env = [] { const char *e = getenv("PDLB200_MM_STREAMK"); return e ? atoi(e) : -1; }();
  const long long ntiles = (long long)grid.x * grid.y;
  const int KT = (int)((p.T + TM_BK - 1) / TM_BK);
  const long long total = ntiles * KT;
  const int sms = sm_count();
  const long long waves = (ntiles + sms - 1) / sms;
  const bool uneven = ntiles * 100 < waves * sms * 97;            // the last wave leaves > 3% of the machine idle
  if (sk_env != 0 && (uneven || sk_env == 1) && total < (1ll << 31) && ntiles < (1ll << 24)) {
    SkPlan k;
    long long G = total / 8;                                      // at least 8 k-tiles (k depth 128) per CTA
    if (G > sms) G = sms;
    if (G < 1) G = 1;
    k.G = (int)G; k.KT = KT; k.ntx = (int)grid.x; k.nty = (int)grid.y; k.total = total;
    const size_t wsb = (size_t)G * 32 * TM_THREADS * sizeof(double);
    char *scr = (char *)scratch(wsb + (size_t)G * sizeof(int), st);
    if (scr) {
      k.ws = (double *)scr; k.flags = (int *)(scr + wsb);
      PDLB200_CUDA_OK(cudaMemsetAsync(k.flags, 0, (size_t)G * sizeof(int), st), E);
      static bool sk_attr = false;
      if (!sk_attr) {
        PDLB200_CUDA_OK(cudaFuncSetAttribute(mm_dmma_tma_sk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TM_SMEM), E);
        sk_attr = true;
      }
      mm_dmma_tma_sk_kernel<<<(unsigned)G, TM_THREADS, TM_SMEM, st>>>(mapA, mapB, p, k);
      note_launch("matmult_dmma_tma");
      PDLB200_CUDA_OK(cudaGetLastError(), E);
      return PDLB200_OK;
    }
  }
  mm_dmma_tma_kernel<<<grid, TM_THREADS, TM_SMEM, st>>>(mapA, mapB, p);
  note_launch("matmult_dmma_tma");
  PDLB200_CUDA_OK(cudaGetLastError(), E);
  return PDLB200_OK;
}

}  // namespace pdlb200
