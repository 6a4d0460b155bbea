// matmult_dmma.cu — double-precision matmult on the FP64 tensor cores (DMMA).
//
// lib/PDL/Primitive.pd:191-264 for double, no BAD values, standard layout (t unit-stride in a,
// w unit-stride in b): C[h][w] = sum_t A[h][t] * B[t][w].  tcgen05.mma has no f64 kind, so the
// FP64 tensor path on sm_100a is mma.sync (SASS: DMMA).  Roofline: FP64 tensor peak;
// 2*T*H*W flop, minimum traffic 8*(H*T + T*W + H*W) bytes.
//
// Default kernel (mm_dmma_ws_kernel): CTA tile 128x128, BK = 16, warp-specialised — 16 consumer
// warps (4 x 4, warp tile 32x32: LDS.64 fragment loads + DMMA only) and 4 producer warps that issue
// every cp.async (LDGSTS.128) of a 4-stage ring, handshaking through per-stage full/empty
// mbarriers (cp.async.mbarrier.arrive.noinc on the producer side).  One producer warp could not
// issue the 2048 16-byte copies of a stage fast enough (consumers sat 17% of the time in
// try_wait, ncu SASS sampling); four do: 35.1 TFLOP/s at 8192^3 = 0.99 of cuBLAS DGEMM, 0.95 of the
// 37.1 TFLOP/s DMMA issue peak measured with tools/ubench/dmma_rate.cu.
// PDLB200_DMMA=sync selects the older 8-warp __syncthreads pipeline (31.2 TFLOP/s).
// Shared-memory rows are padded (A: 20 doubles, B: 132 doubles) so that the 64-bit fragment
// loads of a half-warp hit 16 distinct banks (ncu: no LDS conflicts).
// The MMA adds with fused multiply-add, so results equal the reference's separate
// multiply/add only within tolerance (bit-exact when every product and partial sum is exactly
// representable; tests/test_gpu_parity.py pins both).
#include <cstdlib>
#include <cstring>
#include "matmult.cuh"

namespace pdlb200 {

constexpr int DM_BM = 128, DM_BN = 128, DM_BK = 16, DM_STAGES = 4;
constexpr int DM_LDA = DM_BK + 4;    // 20 doubles: rows 160 B apart (16B-aligned, conflict-free)
constexpr int DM_LDB = DM_BN + 4;    // 132 doubles
constexpr int DM_A_STAGE = DM_BM * DM_LDA;   // doubles
constexpr int DM_B_STAGE = DM_BK * DM_LDB;
constexpr size_t DM_SMEM = (size_t)DM_STAGES * (DM_A_STAGE + DM_B_STAGE) * sizeof(double);

__device__ __forceinline__ void cp_async(void *smem, const void *gmem, int bytes_valid, bool wide) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  if (wide) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(s), "l"(gmem), "r"(bytes_valid));
  else      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" :: "r"(s), "l"(gmem), "r"(bytes_valid));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }

// D(8x8) += A(8x4) * B(4x8).  lane: g = lane>>2, t = lane&3.  a = A[g][t], b = B[t][g], c{0,1} = C[g][2t+{0,1}]
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// D(16x8) += A(16x8) * B(8x8).  a0:(g,t) a1:(g+8,t) a2:(g,t+4) a3:(g+8,t+4); b0:(k=t,n=g) b1:(k=t+4,n=g);
// c0,c1:(g, 2t+{0,1}) c2,c3:(g+8, 2t+{0,1})
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// ALIGNED16: every row start of A, B (and the k/w offsets used) is 16-byte aligned -> 16-byte LDGSTS
// SHAPE: 0 = m8n8k4 instructions, 1 = m16n8k8 (4x fewer, 4x larger tensor-pipe issues)
template <bool ALIGNED16, int SHAPE>
__global__ void __launch_bounds__(256, 1)
mm_dmma_kernel(const __grid_constant__ MmPlan p) {
  extern __shared__ __align__(16) double smem[];
  double *sA = smem;
  double *sB = smem + DM_STAGES * DM_A_STAGE;

  int64_t oa = 0, ob = 0, oc = 0;
  {
    int64_t row = p.z0 + blockIdx.z;
    for (int d = 0; d < p.nd; d++) {
      const int64_t q = (d == p.nd - 1) ? 0 : row / p.dims[d];
      const int64_t i = row - q * p.dims[d];
      oa += i * p.sa[d]; ob += i * p.sb[d]; oc += i * p.sc[d];
      row = q;
    }
  }
  const double *A = reinterpret_cast<const double *>(p.a) + oa;
  const double *B = reinterpret_cast<const double *>(p.b) + ob;
  double *C = reinterpret_cast<double *>(p.c) + oc;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;         // 2 x 4 warps
  const int g = lane >> 2, t4 = lane & 3;
  const int64_t h0 = (int64_t)blockIdx.y * DM_BM, w0 = (int64_t)blockIdx.x * DM_BN;
  const int KT = (int)((p.T + DM_BK - 1) / DM_BK);

  auto load_tile = [&](int kt, int slot) {
    const int64_t k0 = (int64_t)kt * DM_BK;
    double *dA = sA + slot * DM_A_STAGE;
    double *dB = sB + slot * DM_B_STAGE;
    if (ALIGNED16) {
      // A: 128 rows x 8 chunks of 2 doubles; B: 16 rows x 64 chunks
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int c = tid + i * 256;
        const int r = c >> 3, kc = (c & 7) * 2;
        const int64_t h = h0 + r, k = k0 + kc;
        int valid = 0;
        if (h < p.H && k < p.T) valid = (p.T - k >= 2) ? 16 : 8;
        const double *src = valid ? (A + h * p.iah + k) : A;
        cp_async(dA + r * DM_LDA + kc, src, valid, true);
      }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int c = tid + i * 256;
        const int r = c >> 6, wc = (c & 63) * 2;
        const int64_t k = k0 + r, w = w0 + wc;
        int valid = 0;
        if (k < p.T && w < p.W) valid = (p.W - w >= 2) ? 16 : 8;
        const double *src = valid ? (B + k * p.ibt + w) : B;
        cp_async(dB + r * DM_LDB + wc, src, valid, true);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int c = tid + i * 256;
        const int r = c >> 4, kc = c & 15;
        const int64_t h = h0 + r, k = k0 + kc;
        const int valid = (h < p.H && k < p.T) ? 8 : 0;
        const double *src = valid ? (A + h * p.iah + k) : A;
        cp_async(dA + r * DM_LDA + kc, src, valid, false);
      }
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int c = tid + i * 256;
        const int r = c >> 7, wc = c & 127;
        const int64_t k = k0 + r, w = w0 + wc;
        const int valid = (k < p.T && w < p.W) ? 8 : 0;
        const double *src = valid ? (B + k * p.ibt + w) : B;
        cp_async(dB + r * DM_LDB + wc, src, valid, false);
      }
    }
  };

  // accumulators: SHAPE 0: acc[i][j][0..1] = C[8i+g][8j+2t+{0,1}];  SHAPE 1 groups m8 tiles in pairs:
  // acc16[i][j][0..3] = C[16i+g][..], C[16i+g+8][..]
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

#pragma unroll
  for (int s = 0; s < DM_STAGES - 1; s++) {
    if (s < KT) load_tile(s, s);
    cp_commit();
  }

  for (int kt = 0; kt < KT; kt++) {
    cp_wait<DM_STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + DM_STAGES - 1;
      if (nk < KT) load_tile(nk, nk % DM_STAGES);
      cp_commit();
    }
    const double *tA = sA + (kt % DM_STAGES) * DM_A_STAGE + (wm * 64) * DM_LDA;
    const double *tB = sB + (kt % DM_STAGES) * DM_B_STAGE + wn * 32;
    if (SHAPE == 0) {
#pragma unroll
      for (int ks = 0; ks < DM_BK; ks += 4) {
        double af[8], bf[4];
#pragma unroll
        for (int i = 0; i < 8; i++) af[i] = tA[(i * 8 + g) * DM_LDA + ks + t4];
#pragma unroll
        for (int j = 0; j < 4; j++) bf[j] = tB[(ks + t4) * DM_LDB + j * 8 + g];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    } else {
#pragma unroll
      for (int ks = 0; ks < DM_BK; ks += 8) {
        double af[4][4], bf[4][2];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          af[i][0] = tA[(i * 16 + g) * DM_LDA + ks + t4];
          af[i][1] = tA[(i * 16 + g + 8) * DM_LDA + ks + t4];
          af[i][2] = tA[(i * 16 + g) * DM_LDA + ks + t4 + 4];
          af[i][3] = tA[(i * 16 + g + 8) * DM_LDA + ks + t4 + 4];
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          bf[j][0] = tB[(ks + t4) * DM_LDB + j * 8 + g];
          bf[j][1] = tB[(ks + t4 + 4) * DM_LDB + j * 8 + g];
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) {
            double c4[4] = {acc[2 * i][j][0], acc[2 * i][j][1], acc[2 * i + 1][j][0], acc[2 * i + 1][j][1]};
            dmma1688(c4, af[i], bf[j]);
            acc[2 * i][j][0] = c4[0]; acc[2 * i][j][1] = c4[1]; acc[2 * i + 1][j][0] = c4[2]; acc[2 * i + 1][j][1] = c4[3];
          }
      }
    }
  }
  cp_wait<0>();

  // epilogue: each lane owns C[h][w..w+1] pairs
  const bool c_vec = (p.icw == 1) && ((((uintptr_t)C) & 15) == 0) && ((p.ich & 1) == 0);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int64_t h = h0 + wm * 64 + i * 8 + g;
    if (h >= p.H) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int64_t w = w0 + wn * 32 + j * 8 + t4 * 2;
      if (w >= p.W) continue;
      double *dst = C + h * p.ich + w * p.icw;
      if (c_vec && w + 1 < p.W) {
        *reinterpret_cast<double2 *>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
      } else {
        dst[0] = acc[i][j][0];
        if (w + 1 < p.W) dst[p.icw] = acc[i][j][1];
      }
    }
  }
}

template <bool ALIGNED16, int SHAPE>
static int dmma_go(const MmPlan &p, dim3 grid, cudaStream_t s, const Err &E) {
  static bool attr_set = false;
  if (!attr_set) {
    PDLB200_CUDA_OK(cudaFuncSetAttribute(mm_dmma_kernel<ALIGNED16, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DM_SMEM), E);
    attr_set = true;
  }
  mm_dmma_kernel<ALIGNED16, SHAPE><<<grid, 256, DM_SMEM, s>>>(p);
  return PDLB200_OK;
}

// ---- warp-specialised variant ---------------------------------------------------------------
// 8 consumer warps (LDS + DMMA only) + 1 producer warp (all cp.async for both operand tiles),
// connected by per-stage full/empty mbarriers instead of a CTA-wide __syncthreads per k-tile:
// the tensor pipe no longer idles while every warp recomputes load addresses after the barrier.
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" :: "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cp_async(uint64_t *bar) {
  // arrives (without bumping the expected count) once all of this thread's prior cp.async have landed
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" :: "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n"
      :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

constexpr int WS_MAX_STAGES = 6;
constexpr size_t ws_smem(int stages) { return (size_t)stages * (DM_A_STAGE + DM_B_STAGE) * sizeof(double) + 2 * WS_MAX_STAGES * sizeof(uint64_t); }

// NCW consumer warps: 8 -> 2x4 warps of 64x32 tiles, 16 -> 4x4 warps of 32x32 tiles (more warps per
// scheduler to cover the fixed DMMA issue latency)
template <bool ALIGNED16, int NCW, int WS_STAGES, int NPW>
__global__ void __launch_bounds__(NCW * 32 + NPW * 32, 1)
mm_dmma_ws_kernel(const __grid_constant__ MmPlan p) {
  constexpr int MT = (NCW == 8) ? 8 : 4;          // m8 tiles per warp
  extern __shared__ __align__(16) double smem[];
  double *sA = smem;
  double *sB = smem + WS_STAGES * DM_A_STAGE;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + WS_STAGES * (DM_A_STAGE + DM_B_STAGE));
  uint64_t *empty = full + WS_STAGES;

  int64_t oa = 0, ob = 0, oc = 0;
  {
    int64_t row = p.z0 + blockIdx.z;
    for (int d = 0; d < p.nd; d++) {
      const int64_t q = (d == p.nd - 1) ? 0 : row / p.dims[d];
      const int64_t i = row - q * p.dims[d];
      oa += i * p.sa[d]; ob += i * p.sb[d]; oc += i * p.sc[d];
      row = q;
    }
  }
  const double *A = reinterpret_cast<const double *>(p.a) + oa;
  const double *B = reinterpret_cast<const double *>(p.b) + ob;
  double *C = reinterpret_cast<double *>(p.c) + oc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t h0 = (int64_t)blockIdx.y * DM_BM, w0 = (int64_t)blockIdx.x * DM_BN;
  const int KT = (int)((p.T + DM_BK - 1) / DM_BK);

  if (tid == 0) {
    for (int s = 0; s < WS_STAGES; s++) { mbar_init(&full[s], 32 * NPW); mbar_init(&empty[s], NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp >= NCW) {
    // ===== producers: NPW warps share the rows of both tiles of every stage =====
    const int pw = warp - NCW;
    // loop-invariant parts of the aligned copy pattern
    const int a_kc = (lane & 7) * 2;
    const int a_dst = (lane >> 3) * DM_LDA + a_kc;
    const double *a_src = A + (h0 + (lane >> 3)) * p.iah + a_kc;
    unsigned a_mask = 0;
    for (int i = 0; i < 32; i++) if (h0 + (lane >> 3) + 4 * i < p.H) a_mask |= 1u << i;
    const int b_dst = lane * 2;
    const double *b_src = B + w0 + lane * 2;
    int b_bytes[2];
    for (int q = 0; q < 2; q++) {
      const int64_t w = w0 + lane * 2 + 64 * q;
      b_bytes[q] = (w < p.W) ? ((p.W - w >= 2) ? 16 : 8) : 0;
    }
    for (int kt = 0; kt < KT; kt++) {
      const int s = kt % WS_STAGES;
      if (kt >= WS_STAGES) mbar_wait(&empty[s], ((kt / WS_STAGES) - 1) & 1);
      const int64_t k0 = (int64_t)kt * DM_BK;
      double *dA = sA + s * DM_A_STAGE;
      double *dB = sB + s * DM_B_STAGE;
      if (ALIGNED16) {
        // lean issue loop: row validity (A) and column validity (B) are loop-invariant and were
        // folded into a_mask / b_bytes before the k loop; per copy = 1 address add + 1 LDGSTS
        const int64_t k = k0 + a_kc;
        const int kvalid = (k < p.T) ? ((p.T - k >= 2) ? 16 : 8) : 0;
        const double *ga = a_src + k0;
        double *da = dA + a_dst;
#pragma unroll
        for (int j = 0; j < 32 / NPW; j++) {      // A: rows (lane>>3) + 4i, i = pw + NPW*j
          const int i = pw + NPW * j;
          const int valid = ((a_mask >> i) & 1u) ? kvalid : 0;
          cp_async(da + (4 * i) * DM_LDA, valid ? (ga + (int64_t)(4 * i) * p.iah) : A, valid, true);
        }
        const int64_t krem = p.T - k0;             // k-rows of this tile that exist
        const double *gb = b_src + k0 * p.ibt;
        double *db = dB + b_dst;
#pragma unroll
        for (int j = 0; j < 32 / NPW; j++) {      // B: k-row i>>1, chunk lane + 32*(i&1)
          const int i = pw + NPW * j;
          const int r = i >> 1;
          const int valid = (r < krem) ? b_bytes[i & 1] : 0;
          cp_async(db + r * DM_LDB + 64 * (i & 1), valid ? (gb + (int64_t)r * p.ibt + 64 * (i & 1)) : B, valid, true);
        }
      } else {
        const int kc = lane & 15;
        const int64_t k = k0 + kc;
#pragma unroll 8
        for (int i = pw; i < 64; i += NPW) {      // A: rows (lane>>4) + 2i, element lane&15
          const int r = (lane >> 4) + 2 * i;
          const int64_t h = h0 + r;
          const int valid = (h < p.H && k < p.T) ? 8 : 0;
          cp_async(dA + r * DM_LDA + kc, valid ? (A + h * p.iah + k) : A, valid, false);
        }
#pragma unroll 8
        for (int i = pw; i < 64; i += NPW) {      // B: k-row i>>2, element lane + 32*(i&3)
          const int r = i >> 2, wc = lane + 32 * (i & 3);
          const int64_t kk = k0 + r, w = w0 + wc;
          const int valid = (kk < p.T && w < p.W) ? 8 : 0;
          cp_async(dB + r * DM_LDB + wc, valid ? (B + kk * p.ibt + w) : B, valid, false);
        }
      }
      mbar_arrive_cp_async(&full[s]);
    }
    cp_wait<0>();
    return;
  }

  // ===== consumers =====
  const int wm = warp >> 2, wn = warp & 3;
  const int g = lane >> 2, t4 = lane & 3;
  double acc[MT][4][2];
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

  for (int kt = 0; kt < KT; kt++) {
    const int s = kt % WS_STAGES;
    mbar_wait(&full[s], (kt / WS_STAGES) & 1);
    const double *tA = sA + s * DM_A_STAGE + (wm * MT * 8) * DM_LDA;
    const double *tB = sB + s * DM_B_STAGE + wn * 32;
#pragma unroll
    for (int ks = 0; ks < DM_BK; ks += 4) {
      double af[MT], bf[4];
#pragma unroll
      for (int i = 0; i < MT; i++) af[i] = tA[(i * 8 + g) * DM_LDA + ks + t4];
#pragma unroll
      for (int j = 0; j < 4; j++) bf[j] = tB[(ks + t4) * DM_LDB + j * 8 + g];
#pragma unroll
      for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  const bool c_vec = (p.icw == 1) && ((((uintptr_t)C) & 15) == 0) && ((p.ich & 1) == 0);
#pragma unroll
  for (int i = 0; i < MT; i++) {
    const int64_t h = h0 + wm * MT * 8 + i * 8 + g;
    if (h >= p.H) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int64_t w = w0 + wn * 32 + j * 8 + t4 * 2;
      if (w >= p.W) continue;
      double *dst = C + h * p.ich + w * p.icw;
      if (c_vec && w + 1 < p.W) {
        *reinterpret_cast<double2 *>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
      } else {
        dst[0] = acc[i][j][0];
        if (w + 1 < p.W) dst[p.icw] = acc[i][j][1];
      }
    }
  }
}

template <bool ALIGNED16, int NCW, int STAGES, int NPW>
static int dmma_ws_go(const MmPlan &p, dim3 grid, cudaStream_t s, const Err &E) {
  static bool attr_set = false;
  if (!attr_set) {
    PDLB200_CUDA_OK(cudaFuncSetAttribute(mm_dmma_ws_kernel<ALIGNED16, NCW, STAGES, NPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ws_smem(STAGES)), E);
    attr_set = true;
  }
  mm_dmma_ws_kernel<ALIGNED16, NCW, STAGES, NPW><<<grid, NCW * 32 + NPW * 32, ws_smem(STAGES), s>>>(p);
  return PDLB200_OK;
}

static int launch_matmult_dmma_slice(const pdlb200_trans *t, const MmPlan &p, int64_t nz, const Err &E);

int launch_matmult_dmma(const pdlb200_trans *t, const MmPlan &plan, const Err &E) {
  MmPlan p = plan;
  for (p.z0 = 0; p.z0 < p.nbatch || p.z0 == 0; p.z0 += MM_MAXZ) {   // gridDim.z <= 65535: slices
    const int64_t nz = (p.nbatch - p.z0 < MM_MAXZ) ? p.nbatch - p.z0 : MM_MAXZ;
    const int rc = launch_matmult_dmma_slice(t, p, nz, E);
    if (rc) return rc;
    if (p.nbatch <= MM_MAXZ) break;
  }
  return PDLB200_OK;
}

static int launch_matmult_dmma_slice(const pdlb200_trans *t, const MmPlan &p, int64_t nz, const Err &E) {
  // eligibility: unit stride along t in a and along w in b (PDL's default physical layout)
  if (p.T == 0) return PDLB200_EUNSUPPORTED;
  if (!((p.iat == 1 || p.T == 1) && (p.ibw == 1 || p.W == 1))) return PDLB200_EUNSUPPORTED;
  if (p.H * p.W < 64 * 64) return PDLB200_EUNSUPPORTED;   // tiny products: the exact kernel is as fast and bit-exact
  bool aligned = ((((uintptr_t)p.a) | ((uintptr_t)p.b)) & 15) == 0 && (p.iah % 2 == 0) && (p.ibt % 2 == 0) &&
                 p.iat == 1 && p.ibw == 1;
  for (int d = 0; d < p.nd && aligned; d++) if ((p.sa[d] % 2) || (p.sb[d] % 2)) aligned = false;
  dim3 grid((unsigned)((p.W + DM_BN - 1) / DM_BN), (unsigned)((p.H + DM_BM - 1) / DM_BM), (unsigned)nz);
  cudaStream_t s = (cudaStream_t)t->stream;
  const char *sh = getenv("PDLB200_DMMA_SHAPE");
  const int shape = (sh && !strcmp(sh, "884")) ? 0 : 1;
  const char *var = getenv("PDLB200_DMMA");            // "sync" = the __syncthreads pipeline; default warp-specialised
  int rc;
  if (!(var && !strcmp(var, "sync"))) {
    const char *name = "matmult_dmma_ws16_p4";
    if (!aligned) { rc = dmma_ws_go<false, 16, 4, 2>(p, grid, s, E); name = "matmult_dmma_ws16_v8"; }
    else          { rc = dmma_ws_go<true, 16, 4, 4>(p, grid, s, E); }
    if (rc) return rc;
    note_launch(name);
    PDLB200_CUDA_OK(cudaGetLastError(), E);
    return PDLB200_OK;
  }
  if (aligned) rc = shape ? dmma_go<true, 1>(p, grid, s, E) : dmma_go<true, 0>(p, grid, s, E);
  else         rc = shape ? dmma_go<false, 1>(p, grid, s, E) : dmma_go<false, 0>(p, grid, s, E);
  if (rc) return rc;
  note_launch(aligned ? (shape ? "matmult_dmma1688_v16" : "matmult_dmma884_v16") : (shape ? "matmult_dmma1688_v8" : "matmult_dmma884_v8"));
  PDLB200_CUDA_OK(cudaGetLastError(), E);
  return PDLB200_OK;
}

}  // namespace pdlb200
