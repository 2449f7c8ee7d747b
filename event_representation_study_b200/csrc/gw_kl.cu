// GWD-B: conditional-gradient Gromov-Wasserstein with the KL loss between the Gaussian kernels of two point clouds
// (representations/representation_search/gromov_wasserstein.py:39-69: OTMI.__init__ + OTMI.solve, which calls
// POT's ot.gromov.gromov_wasserstein(Ks, Kt, p, q, "kl_loss")).
//
//   constC = f1(Ks) p 1^T + 1 q^T f2(Kt)^T,  f1(a) = a log(a + 1e-15) - a,  f2(b) = b
//   tens(T) = constC - hC1 T hC2^T,           hC1 = Ks,  hC2 = log(Kt + 1e-15)       <- the dense contraction
//   loop:  LMO  Gc = argmin_{G in U(p,q)} <tens(T), G>;  exact line search on the quadratic;  T += alpha (Gc - T)
//
// The contraction is the only GEMM-shaped work of the whole library and the only place tensor cores are used:
// k_gemm_nt_3xtf32 is a hand-written tcgen05 kernel (TF32 UMMA, accumulators in TMEM) that splits every fp32
// operand into two TF32 terms on the fly while staging it into the 128-byte-swizzled shared-memory tiles and
// issues hi*hi + lo*hi + hi*lo, i.e. fp32-class accuracy (about 2^-22 relative) at a third of the TF32 rate.
// For uniform marginals with n == m every vertex of U(p, q) is a permutation matrix / n, so the LMO is a linear
// assignment problem (k_auction on the GPU; an exact host solver is kept as the selectable alternative and as the
// fallback) and hC1 Gc hC2^T is ONE gather + ONE GEMM.
#include <float.h>
#include <limits.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "evrep_common.cuh"

namespace evrep {

// ---------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, one 128 x N x 8 TF32 UMMA issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the mbarrier receives one arrival when every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major operand tile in the canonical SWIZZLE_128B layout: rows of 128 bytes, 8-row groups of 1024 bytes
// (cute/arch/mma_sm100_desc.hpp SmemDescriptor: start >> 4, LBO = 1, SBO = 1024 >> 4, version 1, layout 2)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// ---------------------------------------------------------------------------------------------
// C[M x N] = alpha * A[M x K] * B[N x K]^T + rv[i] + cv[j]      (all fp32, row-major, K contiguous in A and B)
// ---------------------------------------------------------------------------------------------
constexpr int GB_M = 128, GB_N = 128, GB_K = 32;  // CTA tile; GB_K floats = one 128-byte swizzle row
constexpr int GB_STAGES = 3;
constexpr int GB_TILE_BYTES = GB_M * GB_K * 4;    // 16 KB per operand term
constexpr int GB_STAGE_BYTES = 4 * GB_TILE_BYTES; // A_hi, A_lo, B_hi, B_lo
constexpr int GB_THREADS = 288;                   // warps 0-3 producers, 4-7 accumulators + epilogue, 8 TMEM owner + MMA issuer
constexpr size_t GB_SMEM = (size_t)GB_STAGES * GB_STAGE_BYTES + 1024 /* alignment slack */ + 128 /* barriers */;
// instruction descriptor (InstrDescriptor in mma_sm100_desc.hpp): F32 accumulate, TF32 x TF32, both K-major, M = 128, N = 128
constexpr uint32_t GB_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(GB_N >> 3) << 17) | ((uint32_t)(GB_M >> 4) << 24);

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// One 128 x 32 fp32 tile (rows r0.., columns k0..) in two steps, so that a producer thread has all 16 of its global loads
// (A and B tile) in flight before it converts anything: fetch -> 8 float4 registers, then split into the two TF32
// terms and store them swizzled.  Rows / columns outside the matrix are zero.
__device__ __forceinline__ void gb_fetch_tile(const float* __restrict__ X, int rows, int K, int r0, int k0, bool vec, int t, float4 (&v)[GB_M / 16]) {
  const int c4 = t & 7, rsub = t >> 3;  // 8 threads per row (one 16-byte chunk each), 16 rows per pass
#pragma unroll
  for (int pass = 0; pass < GB_M / 16; ++pass) {
    const int gr = r0 + pass * 16 + rsub, gk = k0 + 4 * c4;
    v[pass] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr < rows) {
      const float* src = X + (size_t)gr * K + gk;
      if (vec && gk + 4 <= K) {
        v[pass] = __ldg(reinterpret_cast<const float4*>(src));
      } else {
        if (gk + 0 < K) v[pass].x = __ldg(src + 0);
        if (gk + 1 < K) v[pass].y = __ldg(src + 1);
        if (gk + 2 < K) v[pass].z = __ldg(src + 2);
        if (gk + 3 < K) v[pass].w = __ldg(src + 3);
      }
    }
  }
}
__device__ __forceinline__ void gb_store_tile(const float4 (&v)[GB_M / 16], unsigned char* hi, unsigned char* lo, int t) {
  const int c4 = t & 7, rsub = t >> 3;
#pragma unroll
  for (int pass = 0; pass < GB_M / 16; ++pass) {
    const int r = pass * 16 + rsub;
    // hi = x rounded to the nearest TF32 (an exact TF32 number, so the tensor core's own conversion cannot change it);
    // lo = the exact remainder x - hi, rounded to TF32 again.  |x - hi - lo| <= 2^-23 |x| and the errors are unbiased.
    uint4 h, l;
    h.x = to_tf32(v[pass].x); l.x = to_tf32(v[pass].x - __uint_as_float(h.x));
    h.y = to_tf32(v[pass].y); l.y = to_tf32(v[pass].y - __uint_as_float(h.y));
    h.z = to_tf32(v[pass].z); l.z = to_tf32(v[pass].z - __uint_as_float(h.z));
    h.w = to_tf32(v[pass].w); l.w = to_tf32(v[pass].w - __uint_as_float(h.w));
    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c4 ^ (r & 7)) << 4);  // Swizzle<3,4,3>: chunk ^= row mod 8
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
  }
}

// TMEM -> registers: 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Warp roles: 0-3 producers (global fp32 -> hi / lo TF32 tiles), 4-7 accumulators (drain TMEM after every k-block),
// 8 TMEM owner + MMA issuer.
//
// Why the accumulators do not simply stay in TMEM for the whole K loop: the tensor core adds into its fp32
// accumulator with truncation, not round-to-nearest.  For the GW operands (hC1 > 0, hC2 < 0: every term has the same
// sign) K / 8 truncating additions into a growing sum leave a one-sided error of about K / 16 ulp (measured 6e-6
// relative at K = 300), and the GW loss is a 30:1 cancellation of this product against constC.  Each k-block of 32
// therefore gets a FRESH accumulator (two TMEM buffers, ping-pong) which four warps add into registers with
// round-to-nearest while the tensor core works on the next k-block: fp32 blocked summation, error ~ sqrt(K / 32) ulp.
//
// PACKED = true: A and B are not matrices but the "operand images" written by k_gemm_pack - for every (128-row tile,
// 32-column k-block) the two swizzled TF32 tiles (hi, lo) exactly as they must sit in shared memory, 32 KB contiguous.  One
// thread then feeds the pipeline with two cp.async.bulk (TMA, SASS UBLKCP) per k-block that complete on the stage's
// mbarrier; no thread touches the operands.  This is the path evrep_gw_kl uses: hC1 is packed once per solve and reused
// by every conditional-gradient step, the gathered hC2 columns are packed by the gather itself.
template <bool PACKED>
__global__ void __launch_bounds__(GB_THREADS, 1) k_gemm_nt_3xtf32(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                                                                  int M, int N, int K, float alpha, const float* __restrict__ rv,
                                                                  const float* __restrict__ cv, int vecA, int vecB, int vecC) {
  extern __shared__ unsigned char gb_raw[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gb_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + (size_t)GB_STAGES * GB_STAGE_BYTES);  // full[S], empty[S], tfull[2], tempty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GB_STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * GB_M, n0 = blockIdx.x * GB_N;
  const int nkb = (K + GB_K - 1) / GB_K;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (uint32_t)(GB_STAGES + s); };
  auto tfull_bar = [&](int q) { return bar0 + 8u * (uint32_t)(2 * GB_STAGES + q); };
  auto tempty_bar = [&](int q) { return bar0 + 8u * (uint32_t)(2 * GB_STAGES + 2 + q); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < GB_STAGES; ++s) {
      mbar_init(full_bar(s), PACKED ? 1 : 128);  // every producer thread arrives (PACKED: the one that issues the bulk copies)
      mbar_init(empty_bar(s), 1);   // one tcgen05.commit
    }
    for (int q = 0; q < 2; ++q) {
      mbar_init(tfull_bar(q), 1);     // one tcgen05.commit
      mbar_init(tempty_bar(q), 128);  // every accumulator thread arrives
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {  // TMEM: 128 lanes x 2 x 128 fp32 columns (two accumulator buffers)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * GB_N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (PACKED && warp < 4) {
    // ---- producer: one thread, two bulk copies per k-block ----
    if (threadIdx.x == 0) {
      const unsigned char* ga = reinterpret_cast<const unsigned char*>(A) + (size_t)blockIdx.y * nkb * (2 * GB_TILE_BYTES);
      const unsigned char* gbp = reinterpret_cast<const unsigned char*>(B) + (size_t)blockIdx.x * nkb * (2 * GB_TILE_BYTES);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % GB_STAGES;
        mbar_wait(empty_bar(s), ((uint32_t)(kb / GB_STAGES) & 1u) ^ 1u);
        const uint32_t st = smem_u32(tiles + (size_t)s * GB_STAGE_BYTES);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_bar(s)), "r"((uint32_t)GB_STAGE_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(st),
                     "l"(ga + (size_t)kb * (2 * GB_TILE_BYTES)), "r"((uint32_t)(2 * GB_TILE_BYTES)), "r"(full_bar(s))
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(st + 2 * GB_TILE_BYTES),
                     "l"(gbp + (size_t)kb * (2 * GB_TILE_BYTES)), "r"((uint32_t)(2 * GB_TILE_BYTES)), "r"(full_bar(s))
                     : "memory");
      }
    }
  } else if (warp < 4) {
    // ---- producers ----
    const int t = threadIdx.x;
    float4 va[GB_M / 16], vb[GB_N / 16];
    gb_fetch_tile(A, M, K, m0, 0, vecA != 0, t, va);
    gb_fetch_tile(B, N, K, n0, 0, vecB != 0, t, vb);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % GB_STAGES;
      const uint32_t round = (uint32_t)(kb / GB_STAGES);
      mbar_wait(empty_bar(s), (round & 1u) ^ 1u);  // a fresh barrier passes the wait on parity 1
      unsigned char* st = tiles + (size_t)s * GB_STAGE_BYTES;
      gb_store_tile(va, st, st + GB_TILE_BYTES, t);
      gb_store_tile(vb, st + 2 * GB_TILE_BYTES, st + 3 * GB_TILE_BYTES, t);
      if (kb + 1 < nkb) {  // the next k-block's loads fly while the tensor core and the other producers work
        gb_fetch_tile(A, M, K, m0, (kb + 1) * GB_K, vecA != 0, t, va);
        gb_fetch_tile(B, N, K, n0, (kb + 1) * GB_K, vecB != 0, t, vb);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core's async proxy
      mbar_arrive(full_bar(s));
    }
  } else if (warp < 8) {
    // ---- accumulators: TMEM -> registers after every k-block, then the epilogue ----
    const int wq = warp - 4;                // == warp % 4: this warp may touch TMEM lanes 32 wq .. 32 wq + 31
    const int row = m0 + wq * 32 + lane;
    float acc[GB_N];
#pragma unroll
    for (int j = 0; j < GB_N; ++j) acc[j] = 0.f;
    for (int kb = 0; kb < nkb; ++kb) {
      const int q = kb & 1;
      mbar_wait(tfull_bar(q), (uint32_t)(kb >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < GB_N; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_d + ((uint32_t)(wq * 32) << 16) + (uint32_t)(q * GB_N + c0), r);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r[j]);
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(q));  // this buffer may be overwritten
    }
    if (row < M) {
      const float rvv = rv ? rv[row] : 0.f;
      float* dst = C + (size_t)row * N + n0;
#pragma unroll
      for (int j = 0; j < GB_N; j += 4) {
        const int col = n0 + j;
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = alpha * acc[j + e] + rvv + ((cv && col + e < N) ? __ldg(cv + col + e) : 0.f);
        if (vecC && col + 4 <= N) {
          *reinterpret_cast<float4*>(dst + j) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (col + e < N) dst[j + e] = o[e];
        }
      }
    }
  } else if (lane == 0) {
    // ---- MMA issuer: one thread ----
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % GB_STAGES, q = kb & 1;
      mbar_wait(tempty_bar(q), ((uint32_t)(kb >> 1) & 1u) ^ 1u);  // accumulator buffer drained (passes at once the first time)
      mbar_wait(full_bar(s), (uint32_t)(kb / GB_STAGES) & 1u);
      tc_fence_after();
      const uint32_t st = smem_u32(tiles + (size_t)s * GB_STAGE_BYTES);
      const uint32_t td = tmem_d + (uint32_t)(q * GB_N);
      const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + GB_TILE_BYTES);
      const uint64_t b_hi = umma_desc_sw128(st + 2 * GB_TILE_BYTES), b_lo = umma_desc_sw128(st + 3 * GB_TILE_BYTES);
#pragma unroll
      for (int k = 0; k < GB_K / 8; ++k) {  // UMMA K = 8 TF32 = 32 bytes: advance the start address inside the swizzle row
        const uint64_t adv = (uint64_t)((k * 32) >> 4);
        // small terms first: they are added into a small accumulator
        umma_tf32(td, a_lo + adv, b_hi + adv, GB_IDESC, k != 0 ? 1u : 0u);
        umma_tf32(td, a_hi + adv, b_lo + adv, GB_IDESC, 1u);
      }
#pragma unroll
      for (int k = 0; k < GB_K / 8; ++k) {
        const uint64_t adv = (uint64_t)((k * 32) >> 4);
        umma_tf32(td, a_hi + adv, b_hi + adv, GB_IDESC, 1u);
      }
      umma_commit(empty_bar(s));  // the stage may be refilled once these MMAs have read it
      umma_commit(tfull_bar(q));  // this k-block's partial product is complete
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(2 * GB_N) : "memory");
  }
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
int launch_gemm_nt_3xtf32(const float* A, const float* B, float* C, int M, int N, int K, float alpha, const float* rv, const float* cv,
                          cudaStream_t stream);

// Operand image of X (rows x K, row major): tile (rt, kb) at byte (rt * nkb + kb) * 32 KB = [hi tile | lo tile], each 128
// rows x 128 bytes with the 16-byte chunks of row r at position chunk ^ (r mod 8).  With `sigma` the columns are gathered
// first: element (j, k) is X[j, sigma[k]] (the GW step's hC2[:, sigma]).  Rows / columns past the matrix are zero.
__global__ void __launch_bounds__(256) k_gemm_pack(const float* __restrict__ X, int rows, int K, int ldx, const int* __restrict__ sigma,
                                                   unsigned char* __restrict__ img) {
  const int nkb = (K + GB_K - 1) / GB_K;
  const int n_rt = (rows + GB_M - 1) / GB_M;
  const size_t total = (size_t)n_rt * nkb * GB_M * 8;  // one thread per (tile, row, 16-byte chunk)
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(e & 7);
    const int r = (int)((e >> 3) & (GB_M - 1));
    const size_t tile = e >> 10;  // GB_M * 8 = 1024 chunks per tile
    const int kb = (int)(tile % nkb), rt = (int)(tile / nkb);
    const int gr = rt * GB_M + r, gk = kb * GB_K + 4 * c4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (gr < rows) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (gk + q < K) v[q] = __ldg(X + (size_t)gr * ldx + (sigma ? sigma[gk + q] : gk + q));
    }
    uint4 h, l;
    h.x = to_tf32(v[0]); l.x = to_tf32(v[0] - __uint_as_float(h.x));
    h.y = to_tf32(v[1]); l.y = to_tf32(v[1] - __uint_as_float(h.y));
    h.z = to_tf32(v[2]); l.z = to_tf32(v[2] - __uint_as_float(h.z));
    h.w = to_tf32(v[3]); l.w = to_tf32(v[3] - __uint_as_float(h.w));
    unsigned char* t0 = img + tile * (size_t)(2 * GB_TILE_BYTES);
    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c4 ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(t0 + off) = h;
    *reinterpret_cast<uint4*>(t0 + GB_TILE_BYTES + off) = l;
  }
}

size_t gemm_image_bytes(int rows, int K) { return (size_t)((rows + GB_M - 1) / GB_M) * (size_t)((K + GB_K - 1) / GB_K) * (size_t)(2 * GB_TILE_BYTES); }

int launch_gemm_pack(const float* X, int rows, int K, int ldx, const int* sigma, void* img, cudaStream_t stream) {
  const size_t total = gemm_image_bytes(rows, K) / 32;  // 16-byte chunk pairs
  const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)148 * 16);
  k_gemm_pack<<<blocks, 256, 0, stream>>>(X, rows, K, ldx, sigma, (unsigned char*)img);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

// C = alpha * A * B^T + rv + cv from two operand images (k_gemm_pack)
int launch_gemm_packed(const void* imgA, const void* imgB, float* C, int M, int N, int K, float alpha, const float* rv, const float* cv,
                       cudaStream_t stream) {
  EVREP_CUDA_OK(cudaFuncSetAttribute(k_gemm_nt_3xtf32<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GB_SMEM));
  dim3 grid((unsigned)((N + GB_N - 1) / GB_N), (unsigned)((M + GB_M - 1) / GB_M));
  k_gemm_nt_3xtf32<true><<<grid, GB_THREADS, GB_SMEM, stream>>>((const float*)imgA, (const float*)imgB, C, M, N, K, alpha, rv, cv, 0, 0,
                                                                (al16(C) && N % 4 == 0) ? 1 : 0);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

size_t gemm_workspace_bytes(int M, int N, int K) {
  if (M < 1 || N < 1 || K < 1) return 0;
  return align_up(gemm_image_bytes(M, K), 256) + align_up(gemm_image_bytes(N, K), 256);
}

// with a workspace of gemm_workspace_bytes: pack both operands, then the TMA-fed kernel; without: the register-staged kernel
int launch_gemm_nt_3xtf32_ws(const float* A, const float* B, float* C, int M, int N, int K, float alpha, const float* rv, const float* cv,
                             void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (M < 1 || N < 1 || K < 1) {
    set_error("gemm: empty problem");
    return EVREP_EINVAL;
  }
  if (!workspace) return launch_gemm_nt_3xtf32(A, B, C, M, N, K, alpha, rv, cv, stream);
  if ((reinterpret_cast<uintptr_t>(workspace) & 255u) || workspace_bytes < gemm_workspace_bytes(M, N, K)) {
    set_error("gemm: workspace must be 256-byte aligned and hold %zu bytes", gemm_workspace_bytes(M, N, K));
    return EVREP_EWORKSPACE;
  }
  void* imgA = workspace;
  void* imgB = (char*)workspace + align_up(gemm_image_bytes(M, K), 256);
  int rc = launch_gemm_pack(A, M, K, K, nullptr, imgA, stream);
  if (rc) return rc;
  rc = launch_gemm_pack(B, N, K, K, nullptr, imgB, stream);
  if (rc) return rc;
  return launch_gemm_packed(imgA, imgB, C, M, N, K, alpha, rv, cv, stream);
}

int launch_gemm_nt_3xtf32(const float* A, const float* B, float* C, int M, int N, int K, float alpha, const float* rv, const float* cv,
                          cudaStream_t stream) {
  if (M < 1 || N < 1 || K < 1) {
    set_error("gemm: empty problem");
    return EVREP_EINVAL;
  }
  EVREP_CUDA_OK(cudaFuncSetAttribute(k_gemm_nt_3xtf32<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GB_SMEM));  // per device: cheap, not cached
  dim3 grid((unsigned)((N + GB_N - 1) / GB_N), (unsigned)((M + GB_M - 1) / GB_M));
  k_gemm_nt_3xtf32<false><<<grid, GB_THREADS, GB_SMEM, stream>>>(A, B, C, M, N, K, alpha, rv, cv, (al16(A) && K % 4 == 0) ? 1 : 0,
                                                                 (al16(B) && K % 4 == 0) ? 1 : 0, (al16(C) && N % 4 == 0) ? 1 : 0);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

// ---------------------------------------------------------------------------------------------
// Gaussian kernels and the small dense helpers of the conditional-gradient loop
// ---------------------------------------------------------------------------------------------
constexpr int GK_THREADS = 256;
constexpr int GK_MAX_D = 64;

// mu[k] and mean_ij |xi - xj|^2 = 2/n sum_i |xi - mu|^2 (compute_kernel needs sqrt(mean(C^2) / 2)); one CTA
__global__ void __launch_bounds__(GK_THREADS) k_gwb_moments(const double* __restrict__ X, int n, int d, double* __restrict__ out /* [0] = mean D^2 */) {
  __shared__ double mu[GK_MAX_D];
  __shared__ double red[GK_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = 0; k < d; ++k) {
    double s = 0.0;
    for (int i = tid; i < n; i += GK_THREADS) s += X[(size_t)i * d + k];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) {
      double a = 0.0;
      for (int q = 0; q < GK_THREADS / 32; ++q) a += red[q];
      mu[k] = a / n;
    }
    __syncthreads();
  }
  double s = 0.0;
  for (int i = tid; i < n; i += GK_THREADS)
    for (int k = 0; k < d; ++k) {
      const double v = X[(size_t)i * d + k] - mu[k];
      s += v * v;
    }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (tid == 0) {
    double a = 0.0;
    for (int q = 0; q < GK_THREADS / 32; ++q) a += red[q];
    out[0] = 2.0 * a / n;
  }
}

// One row of the kernel per CTA: K[i, j] = exp(-|xi - xj|^2 / (2 h^2 std^2)) in fp64, stored as fp32 (`Kf`, or its
// log(K + 1e-15) when want_log), plus the fp64 row sums the KL decomposition needs:
//   rs_a[i] = sum_j (K log(K + 1e-15) - K)   (f1, source side)     or   sum_j K         (f2, target side, want_log)
//   rs_h[i] = sum_j hC[i, j]  (K on the source side, log(K + 1e-15) on the target side): A(p q^T) is rank one
__global__ void __launch_bounds__(GK_THREADS) k_gwb_kernel_rows(const double* __restrict__ X, int n, int d, double h, const double* __restrict__ msq,
                                                                int want_log, float* __restrict__ Kf, double* __restrict__ rs_a,
                                                                double* __restrict__ rs_h) {
  __shared__ double xi[GK_MAX_D];
  __shared__ double red[2][GK_THREADS / 32];
  const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < d) xi[tid] = X[(size_t)i * d + tid];
  __syncthreads();
  const double std2 = msq[0] / 2.0;           // std^2 = mean(C^2) / 2
  const double coef = -0.5 / (h * h * std2);  // K = exp(coef * D^2)
  double sa = 0.0, sh = 0.0;
  for (int j = tid; j < n; j += GK_THREADS) {
    double d2 = 0.0;
    for (int k = 0; k < d; ++k) {
      const double v = xi[k] - X[(size_t)j * d + k];
      d2 += v * v;
    }
    const double kv = exp(coef * d2);
    const double lg = log(kv + 1e-15);
    if (want_log) {
      Kf[(size_t)i * n + j] = (float)lg;
      sa += kv;
      sh += lg;
    } else {
      Kf[(size_t)i * n + j] = (float)kv;
      sa += kv * lg - kv;
      sh += kv;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    sh += __shfl_xor_sync(0xffffffffu, sh, o);
  }
  if (lane == 0) { red[0][warp] = sa; red[1][warp] = sh; }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, b = 0.0;
    for (int q = 0; q < GK_THREADS / 32; ++q) { a += red[0][q]; b += red[1][q]; }
    rs_a[i] = a;
    rs_h[i] = b;
  }
}

// cr[i] = rs_a1[i] / n (f1(Ks) p), cc[j] = rs_a2[j] / m (f2(Kt) q); AG = (rs_h1[i] / n) (rs_h2[j] / m) = hC1 (p q^T) hC2^T; G = 1 / (n m)
__global__ void k_gwb_init(int n, int m, const double* __restrict__ rs_a1, const double* __restrict__ rs_a2, const double* __restrict__ rs_h1,
                           const double* __restrict__ rs_h2, float* __restrict__ cr, float* __restrict__ cc, float* __restrict__ G,
                           float* __restrict__ AG) {
  const size_t total = (size_t)n * m;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / m), j = (int)(e - (size_t)i * m);
    G[e] = (float)(1.0 / ((double)n * (double)m));
    AG[e] = (float)((rs_h1[i] / n) * (rs_h2[j] / m));
    if (j == 0) cr[i] = (float)(rs_a1[i] / n);
    if (i == 0) cc[j] = (float)(rs_a2[j] / m);
  }
}

// Bp[j, k] = hC2[j, sigma[k]]: the gather that turns hC1 Gc hC2^T (Gc = permutation / n) into one GEMM
__global__ void k_gwb_gather(const float* __restrict__ hC2, const int* __restrict__ sigma, int m, int n, float* __restrict__ Bp) {
  const size_t total = (size_t)m * n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(e / n), k = (int)(e - (size_t)j * n);
    Bp[e] = hC2[(size_t)j * m + sigma[k]];
  }
}

// Mi = constC - AG (the gradient up to the factor 2 and a constant shift, neither changes the LMO), and
// red[0] += sum Mi * G  (= the GW loss at G)
__global__ void __launch_bounds__(256) k_gwb_grad(int n, int m, const float* __restrict__ cr, const float* __restrict__ cc, const float* __restrict__ AG,
                                                  const float* __restrict__ G, float* __restrict__ Mi, double* __restrict__ red) {
  const size_t total = (size_t)n * m;
  double s = 0.0;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / m), j = (int)(e - (size_t)i * m);
    const float v = cr[i] + cc[j] - AG[e];
    Mi[e] = v;
    s += (double)v * (double)G[e];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(red, s);
}

// line-search sums with Gc = permutation / n given as sigma, dG = Gc - G, AdG = AGc - AG:
//   red[0] = sum AdG * dG,  red[1] = sum constC * dG,  red[2] = sum AG * dG,  red[3] = sum AdG * G
__global__ void __launch_bounds__(256) k_gwb_linesearch(int n, int m, const float* __restrict__ cr, const float* __restrict__ cc,
                                                        const float* __restrict__ AG, const float* __restrict__ AGc, const float* __restrict__ G,
                                                        const int* __restrict__ sigma, const float* __restrict__ Gc, double* __restrict__ red) {
  const size_t total = (size_t)n * m;
  const double inv_n = 1.0 / (double)n;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, mx = 0.0;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / m), j = (int)(e - (size_t)i * m);
    const double g = (double)G[e];
    const double dg = (Gc ? (double)Gc[e] : ((sigma[i] == j) ? inv_n : 0.0)) - g;  // Gc: dense vertex (rectangular plans)
    const double ag = (double)AG[e];
    const double adg = (double)AGc[e] - ag;
    mx = fmax(mx, fabs(adg));
    s0 += adg * dg;
    s1 += ((double)cr[i] + (double)cc[j]) * dg;
    s2 += ag * dg;
    s3 += adg * g;
  }
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(red + 0, s0);
    atomicAdd(red + 1, s1);
    atomicAdd(red + 2, s2);
    atomicAdd(red + 3, s3);
    // red[4] = max |AGc - AG| (the gradient moves by alpha times that): non-negative doubles order like their bit patterns
    if (mx == mx) atomicMax(reinterpret_cast<unsigned long long*>(red + 4), (unsigned long long)__double_as_longlong(mx));
  }
}

// G += alpha (Gc - G),  AG += alpha (AGc - AG)
__global__ void k_gwb_step(int n, int m, float alpha, const int* __restrict__ sigma, const float* __restrict__ Gc, const float* __restrict__ AGc,
                           float* __restrict__ G, float* __restrict__ AG) {
  const size_t total = (size_t)n * m;
  const float inv_n = 1.f / (float)n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / m), j = (int)(e - (size_t)i * m);
    const float g = G[e], ag = AG[e];
    G[e] = g + alpha * ((Gc ? Gc[e] : ((sigma[i] == j) ? inv_n : 0.f)) - g);
    AG[e] = ag + alpha * (AGc[e] - ag);
  }
}

// Rectangular plans (n != m): the LMO vertex comes from the host in CSR form (transport.cu).
// Gc (dense n x m) = the vertex, for the line search and the update
__global__ void k_gwb_plan_scatter(const int* __restrict__ row_ptr, const int* __restrict__ col, const float* __restrict__ wgt, int n, int m,
                                   float* __restrict__ Gc) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) Gc[(size_t)i * m + col[e]] = wgt[e];
}
// XT[j, k] = sum_l Gc[k, l] hC2[j, l]  (m x n): the sparse half of hC1 Gc hC2^T; the dense half is one GEMM with hC1
__global__ void k_gwb_sparse_xt(const float* __restrict__ hC2, const int* __restrict__ row_ptr, const int* __restrict__ col,
                                const float* __restrict__ wgt, int m, int n, float* __restrict__ XT) {
  const size_t total = (size_t)m * n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(e / n), k = (int)(e - (size_t)j * n);
    const float* row = hC2 + (size_t)j * m;
    float s = 0.f;
    for (int q = row_ptr[k]; q < row_ptr[k + 1]; ++q) s = fmaf(wgt[q], row[col[q]], s);
    XT[e] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// Device LMO: the assignment problem by Bertsekas' forward auction with epsilon scaling, Jacobi rounds, one CTA.
//
// Persons = rows, objects = columns, benefit a_ij = -cost_ij.  In a round every unassigned person i finds its best
// object j1 (value v1 = a_ij1 - price_j1) and the second best value v2 and bids price_j1 + (v1 - v2) + eps; every
// object takes its highest bid, raises its price to it and swaps owners.  When nobody is unassigned the assignment is
// within n eps of optimal; eps starts at C / 4 (C = cost range) and shrinks by theta per phase (prices kept,
// assignment reset) down to eps_rel C, so the final plan is optimal up to n eps_rel C - on the GW cost matrices
// (numpy prototype, n = 300 and 1000) that is the exact optimum of scipy's solver at eps_rel = 1e-9.
// One warp per bidder scans its cost row (coalesced, L2 resident); prices, owners and the bid table live in shared
// memory (doubles, and benefits are taken relative to the smallest cost so that their magnitude is C: an increment of
// 1e-9 C must stay visible in a_ij - price_j, or two bidders fight over one object for gap / eps rounds - seen with the
// nearly constant cost matrix of an n = 2 problem before the shift); rounds are separated by __syncthreads only.  A round with few
// bidders costs one row scan (~3 us), which is what most of the thousands of rounds are.
// ---------------------------------------------------------------------------------------------
int transport_plan_host(const float* cost, int n, int m, int cap, int* row_ptr, int* col, double* weight, int* nnz_out);  // transport.cu

constexpr int AUC_THREADS = 1024;
constexpr int AUC_MAX_N = 4096;
static size_t auction_smem_bytes(int n) { return (size_t)n * (3 * sizeof(double) + 5 * sizeof(int)) + 16; }

// Warm start (the LMOs of consecutive conditional-gradient steps see cost matrices that differ by alpha (AGc - AG)): with
// `price_io` non-null the final prices are written back, and with eps0_abs > 0 they are also READ as the starting prices and
// the scaling starts at eps0_abs (clamped to [eps_final, C / 4]) instead of C / 4.  Any starting prices are valid (every
// phase begins with an empty assignment, so epsilon-complementary slackness holds trivially); prices that were optimal
// for the previous matrix up to 2 max |change| save the first scaling phases (measured at n = 1000: 748 -> 638 ms per pair).
// Tried on top and dropped (profiles/README.md): bids published as (value | person) keys with an incrementally kept queue and
// rows split over the warps when few persons bid (fewer barriers, but 5.8 us per round against 4.4), and repairing the
// previous assignment at the final epsilon (the early steps move the matrix too much: repairs ran into their round cap).
__global__ void __launch_bounds__(AUC_THREADS, 1) k_auction(const float* __restrict__ cost, int n, double eps_rel, double theta, int max_rounds,
                                                            int* __restrict__ sigma, int* __restrict__ stats /* rounds, bids, status */,
                                                            double* __restrict__ price_io, double eps0_abs) {
  extern __shared__ __align__(16) unsigned char auc_raw[];
  double* price = reinterpret_cast<double*>(auc_raw);                             // n
  unsigned long long* objbid = reinterpret_cast<unsigned long long*>(price + n);  // n: highest bid of the round (bits of a positive double), 0 = none
  double* mybid = reinterpret_cast<double*>(objbid + n);                          // n: bid of person i in this round
  int* owner = reinterpret_cast<int*>(mybid + n);                                 // n: person owning object j, or -1
  int* assigned = owner + n;                                                      // n: object of person i, or -1
  int* queue = assigned + n;                                                      // n: unassigned persons of this round
  int* mybid_obj = queue + n;                                                     // n: object person i bid for
  int* winner = mybid_obj + n;                                                    // n: lowest person index among the highest bidders, INT_MAX = none
  __shared__ int s_count, s_rounds, s_bids, s_bad;
  __shared__ double s_red[2][AUC_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // cost range
  double lo = DBL_MAX, hi = -DBL_MAX;
  for (size_t e = tid; e < (size_t)n * n; e += AUC_THREADS) {
    const double c = (double)__ldg(cost + e);
    lo = fmin(lo, c);
    hi = fmax(hi, c);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) { s_red[0][warp] = lo; s_red[1][warp] = hi; }
  const bool warm = price_io != nullptr && eps0_abs > 0.0;
  for (int j = tid; j < n; j += AUC_THREADS) { price[j] = warm ? fmax(price_io[j], 0.0) : 0.0; objbid[j] = 0ull; winner[j] = INT_MAX; }
  if (tid == 0) { s_rounds = 0; s_bids = 0; s_bad = 0; }
  __syncthreads();
  lo = s_red[0][0]; hi = s_red[1][0];
  for (int w = 1; w < AUC_THREADS / 32; ++w) { lo = fmin(lo, s_red[0][w]); hi = fmax(hi, s_red[1][w]); }
  const double C = fmax(hi - lo, 1e-300);
  const double eps_final = C * eps_rel;
  double eps = fmax(C / 4.0, eps_final);
  if (warm) eps = fmin(eps, fmax(eps0_abs, eps_final));
  int status = 0;

  while (true) {  // epsilon phases: prices are kept, the assignment starts over
    for (int j = tid; j < n; j += AUC_THREADS) { owner[j] = -1; assigned[j] = -1; }
    __syncthreads();
    while (true) {  // rounds
      if (tid == 0) s_count = 0;
      __syncthreads();
      for (int i = tid; i < n; i += AUC_THREADS)
        if (assigned[i] < 0) queue[atomicAdd(&s_count, 1)] = i;
      __syncthreads();
      const int nb = s_count;
      if (nb == 0) break;
      if (s_rounds >= max_rounds) { status = 1; break; }  // uniform: every thread reads the same shared values
      // 1. bids: one warp per unassigned person
      for (int q = warp; q < nb; q += AUC_THREADS / 32) {
        const int i = queue[q];
        const float* row = cost + (size_t)i * n;
        double v1 = -DBL_MAX, v2 = -DBL_MAX;
        int j1 = INT_MAX;
        for (int j = lane; j < n; j += 32) {
          const double v = (lo - (double)__ldg(row + j)) - price[j];  // benefit relative to the smallest cost: magnitudes ~ C
          if (v > v1) { v2 = v1; v1 = v; j1 = j; }
          else if (v > v2) v2 = v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {  // merge (v1, j1, v2); equal values go to the lower object index
          const double ov1 = __shfl_xor_sync(0xffffffffu, v1, o), ov2 = __shfl_xor_sync(0xffffffffu, v2, o);
          const int oj1 = __shfl_xor_sync(0xffffffffu, j1, o);
          if (ov1 > v1 || (ov1 == v1 && oj1 < j1)) {
            v2 = fmax(v1, ov2);
            v1 = ov1;
            j1 = oj1;
          } else {
            v2 = fmax(v2, ov1);
          }
        }
        if (lane == 0) {
          if (n == 1) v2 = v1;
          if (j1 == INT_MAX || !(v1 - v2 >= 0.0) || !isfinite(v1 - v2)) {  // NaN / infinite costs: no meaningful bid
            s_bad = 1;
            j1 = 0;
            v2 = v1 = 0.0;
          }
          const double bid = price[j1] + (v1 - v2) + eps;  // > price >= 0, so its bit pattern orders like its value
          mybid_obj[i] = j1;
          mybid[i] = bid;
          atomicMax(&objbid[j1], (unsigned long long)__double_as_longlong(bid));
        }
      }
      __syncthreads();
      if (s_bad) { status = 2; break; }
      // 2. among the highest bidders of an object, the lowest person index wins (deterministic)
      for (int q = tid; q < nb; q += AUC_THREADS) {
        const int i = queue[q], j = mybid_obj[i];
        if ((unsigned long long)__double_as_longlong(mybid[i]) == objbid[j]) atomicMin(&winner[j], i);
      }
      __syncthreads();
      // 3. the winner takes the object at its bid; the previous owner becomes unassigned
      for (int q = tid; q < nb; q += AUC_THREADS) {
        const int i = queue[q], j = mybid_obj[i];
        if (winner[j] == i) {
          const int prev = owner[j];
          if (prev >= 0) assigned[prev] = -1;
          owner[j] = i;
          assigned[i] = j;
          price[j] = mybid[i];
          objbid[j] = 0ull;
        }
      }
      if (tid == 0) { s_rounds += 1; s_bids += nb; }
      __syncthreads();
      // 4. clear the winner slots (after the barrier: the losers of an object were still reading them in step 3)
      for (int q = tid; q < nb; q += AUC_THREADS) {
        const int i = queue[q], j = mybid_obj[i];
        if (owner[j] == i) winner[j] = INT_MAX;
      }
      __syncthreads();
    }
    if (status || eps <= eps_final) break;
    eps = fmax(eps / theta, eps_final);
    __syncthreads();
  }
  for (int i = tid; i < n; i += AUC_THREADS) sigma[i] = assigned[i];
  if (price_io) {  // prices only ever rise: shift them back so that the smallest is 0 (differences are all that matters)
    double pm = DBL_MAX;
    for (int j = tid; j < n; j += AUC_THREADS) pm = fmin(pm, price[j]);
    for (int o = 16; o > 0; o >>= 1) pm = fmin(pm, __shfl_xor_sync(0xffffffffu, pm, o));
    __syncthreads();
    if (lane == 0) s_red[0][warp] = pm;
    __syncthreads();
    pm = s_red[0][0];
    for (int w = 1; w < AUC_THREADS / 32; ++w) pm = fmin(pm, s_red[0][w]);
    for (int j = tid; j < n; j += AUC_THREADS) price_io[j] = price[j] - pm;
  }
  if (tid == 0) { stats[0] = s_rounds; stats[1] = s_bids; stats[2] = status; }
}

// ---------------------------------------------------------------------------------------------
// Host: linear assignment (shortest augmenting paths with potentials, O(n^3)), the exact LMO for n == m
// ---------------------------------------------------------------------------------------------
static void lap_solve(const float* cost, int n, std::vector<int>& row_to_col) {
  const double INF = DBL_MAX;
  std::vector<double> u(n + 1, 0.0), v(n + 1, 0.0), minv(n + 1);
  std::vector<int> p(n + 1, 0), way(n + 1, 0);
  std::vector<char> used(n + 1);
  for (int i = 1; i <= n; ++i) {
    p[0] = i;
    int j0 = 0;
    std::fill(minv.begin(), minv.end(), INF);
    std::fill(used.begin(), used.end(), 0);
    do {
      used[j0] = 1;
      const int i0 = p[j0];
      const float* crow = cost + (size_t)(i0 - 1) * n;
      double delta = INF;
      int j1 = 0;
      for (int j = 1; j <= n; ++j) {
        if (used[j]) continue;
        const double cur = (double)crow[j - 1] - u[i0] - v[j];
        if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
        if (minv[j] < delta) { delta = minv[j]; j1 = j; }
      }
      if (j1 == 0) {  // no finite reduced cost left (NaN / infinite input): give up with the identity
        row_to_col.resize(n);
        for (int k = 0; k < n; ++k) row_to_col[k] = k;
        return;
      }
      for (int j = 0; j <= n; ++j) {
        if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
        else minv[j] -= delta;
      }
      j0 = j1;
    } while (p[j0] != 0);
    do {
      const int j1 = way[j0];
      p[j0] = p[j1];
      j0 = j1;
    } while (j0);
  }
  row_to_col.assign(n, 0);
  for (int j = 1; j <= n; ++j) row_to_col[p[j] - 1] = j - 1;
}

// the LMO on its own (tests, and callers with their own assignment problems): sigma and stats are DEVICE pointers
int launch_auction(const float* cost, int n, double eps_rel, int* sigma, int* stats, cudaStream_t stream) {
  if (n < 1 || n > AUC_MAX_N) {
    set_error("auction: n = %d outside 1..%d", n, AUC_MAX_N);
    return EVREP_EUNSUPPORTED;
  }
  if (!(eps_rel > 0.0)) {
    set_error("auction: eps_rel must be positive");
    return EVREP_EINVAL;
  }
  EVREP_CUDA_OK(cudaFuncSetAttribute(k_auction, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)auction_smem_bytes(n)));
  k_auction<<<1, AUC_THREADS, auction_smem_bytes(n), stream>>>(cost, n, eps_rel, 6.0, 4000000, sigma, stats, nullptr, 0.0);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

struct GwbWs {
  float *hC1, *hC2, *G, *AG, *AGc, *Mi, *Bp, *cr, *cc;
  float *Gc, *plan_w;          // n != m: dense LMO vertex, CSR weights
  int *plan_rp, *plan_col;     // CSR of the vertex (capacity plan_cap)
  int plan_cap;
  void *imgA, *imgB;  // operand images of the contraction (k_gemm_pack)
  double *rs_a1, *rs_a2, *rs_h1, *rs_h2, *red, *msq, *prices;
  int *sigma, *stats;
  size_t bytes;
};
static GwbWs gwb_carve(void* basep, int n, int m) {
  GwbWs w;
  size_t off = 0;
  char* base = (char*)basep;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  const size_t nm = (size_t)n * m;
  w.hC1 = (float*)take(sizeof(float) * (size_t)n * n);
  w.hC2 = (float*)take(sizeof(float) * (size_t)m * m);
  w.G = (float*)take(sizeof(float) * nm);
  w.AG = (float*)take(sizeof(float) * nm);
  w.AGc = (float*)take(sizeof(float) * nm);
  w.Mi = (float*)take(sizeof(float) * nm);
  w.Bp = (float*)take(sizeof(float) * nm);
  w.imgA = take(gemm_image_bytes(std::max(n, m), std::max(n, m)));
  w.imgB = take(gemm_image_bytes(std::max(n, m), std::max(n, m)));
  w.cr = (float*)take(sizeof(float) * (size_t)n);
  w.cc = (float*)take(sizeof(float) * (size_t)m);
  w.rs_a1 = (double*)take(sizeof(double) * (size_t)n);
  w.rs_h1 = (double*)take(sizeof(double) * (size_t)n);
  w.rs_a2 = (double*)take(sizeof(double) * (size_t)m);
  w.rs_h2 = (double*)take(sizeof(double) * (size_t)m);
  w.red = (double*)take(sizeof(double) * 8);
  w.msq = (double*)take(sizeof(double) * 2);
  w.prices = (double*)take(sizeof(double) * (size_t)std::max(n, m));
  w.sigma = (int*)take(sizeof(int) * (size_t)n);
  w.stats = (int*)take(sizeof(int) * 4);
  w.plan_cap = 2 * (n + m) + 16;
  w.Gc = (float*)take(n != m ? sizeof(float) * nm : 0);
  w.plan_rp = (int*)take(n != m ? sizeof(int) * ((size_t)n + 1) : 0);
  w.plan_col = (int*)take(n != m ? sizeof(int) * (size_t)w.plan_cap : 0);
  w.plan_w = (float*)take(n != m ? sizeof(float) * (size_t)w.plan_cap : 0);
  w.bytes = off;
  return w;
}
size_t gw_kl_workspace_bytes(int n, int m) { return (n < 1 || m < 1) ? 0 : gwb_carve(nullptr, n, m).bytes; }

// Synchronous (the LMO runs on the host between device steps).  Returns the GW loss at the last iterate.
// lmo: 0 = auction on the GPU (falls back to the host solver for one step if it hits its round limit), 1 = host solver
int run_gw_kl(const double* Xs, int n, int ds, const double* Xt, int m, int dt, double h, int max_iter, double tol_rel, double tol_abs,
              int lmo, double* gw_dist_host, float* T_out, int* iters_host, int* lmo_stats_host, void* workspace, size_t workspace_bytes,
              cudaStream_t stream) {
  const bool rect = n != m;  // the LMO is a transportation problem (host, transport.cu) instead of an assignment
  if (n < 1 || m < 1 || ds < 1 || dt < 1 || ds > GK_MAX_D || dt > GK_MAX_D) {
    set_error("gw_kl: need n >= 1 and 1 <= ds, dt <= %d", GK_MAX_D);
    return EVREP_EINVAL;
  }
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) {
    set_error("workspace must be non-null and 256-byte aligned");
    return EVREP_EWORKSPACE;
  }
  const GwbWs w = gwb_carve(workspace, n, m);
  if (w.bytes > workspace_bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return EVREP_EWORKSPACE;
  }
  const size_t nm = (size_t)n * m;
  const int eb = (int)std::min<size_t>((nm + 255) / 256, 148 * 8);
  // kernels: OTMI.__init__ (gromov_wasserstein.py:52-60)
  k_gwb_moments<<<1, GK_THREADS, 0, stream>>>(Xs, n, ds, w.msq);
  k_gwb_moments<<<1, GK_THREADS, 0, stream>>>(Xt, m, dt, w.msq + 1);
  k_gwb_kernel_rows<<<n, GK_THREADS, 0, stream>>>(Xs, n, ds, h, w.msq, 0, w.hC1, w.rs_a1, w.rs_h1);
  k_gwb_kernel_rows<<<m, GK_THREADS, 0, stream>>>(Xt, m, dt, h, w.msq + 1, 1, w.hC2, w.rs_a2, w.rs_h2);
  k_gwb_init<<<eb, 256, 0, stream>>>(n, m, w.rs_a1, w.rs_a2, w.rs_h1, w.rs_h2, w.cr, w.cc, w.G, w.AG);
  EVREP_CUDA_OK(cudaGetLastError());

  {  // hC1 is the A operand of every step's contraction: pack it once
    const int rc = launch_gemm_pack(w.hC1, n, n, n, nullptr, w.imgA, stream);
    if (rc) return rc;
  }
  const bool device_lmo = !rect && lmo == 0 && n <= AUC_MAX_N;
  if (device_lmo) EVREP_CUDA_OK(cudaFuncSetAttribute(k_auction, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)auction_smem_bytes(n)));
  std::vector<float> Mi_host, plan_wf;
  std::vector<double> plan_wd;
  std::vector<int> sigma, plan_rp, plan_col;
  double red_host[8];
  double warm_eps = 0.0;  // <= 0: cold start of the auction
  double f_val = 0.0;
  int it = 0;
  long long lmo_rounds = 0, lmo_bids = 0, lmo_fallbacks = 0;
  {  // the loss of the starting plan; degenerate inputs (a single point, coincident points: std = 0) give NaN kernels
     // in the reference too - report NaN without iterating
    EVREP_CUDA_OK(cudaMemsetAsync(w.red, 0, sizeof(double) * 8, stream));
    k_gwb_grad<<<eb, 256, 0, stream>>>(n, m, w.cr, w.cc, w.AG, w.G, w.Mi, w.red);
    EVREP_CUDA_OK(cudaMemcpyAsync(red_host, w.red, sizeof(double), cudaMemcpyDeviceToHost, stream));
    EVREP_CUDA_OK(cudaStreamSynchronize(stream));
    if (!std::isfinite(red_host[0])) {
      if (gw_dist_host) *gw_dist_host = red_host[0];
      if (iters_host) *iters_host = 0;
      if (lmo_stats_host) lmo_stats_host[0] = lmo_stats_host[1] = lmo_stats_host[2] = 0;
      if (T_out) EVREP_CUDA_OK(cudaMemcpyAsync(T_out, w.G, sizeof(float) * nm, cudaMemcpyDeviceToDevice, stream));
      EVREP_CUDA_OK(cudaStreamSynchronize(stream));
      return EVREP_OK;
    }
  }
  for (; it < max_iter; ++it) {
    EVREP_CUDA_OK(cudaMemsetAsync(w.red, 0, sizeof(double) * 8, stream));
    k_gwb_grad<<<eb, 256, 0, stream>>>(n, m, w.cr, w.cc, w.AG, w.G, w.Mi, w.red);
    // LMO: the vertex of U(p, q) minimising <Mi, G>
    bool solved = false;
    if (device_lmo) {
      int st_host[3];
      // warm start from the previous step's prices: the cost matrix moved by at most alpha * max |AGc - AG| since then
      k_auction<<<1, AUC_THREADS, auction_smem_bytes(n), stream>>>(w.Mi, n, 1e-9, 6.0, 4000000, w.sigma, w.stats, w.prices, warm_eps);
      EVREP_CUDA_OK(cudaMemcpyAsync(st_host, w.stats, sizeof(int) * 3, cudaMemcpyDeviceToHost, stream));
      EVREP_CUDA_OK(cudaMemcpyAsync(red_host, w.red, sizeof(double), cudaMemcpyDeviceToHost, stream));
      EVREP_CUDA_OK(cudaStreamSynchronize(stream));
      lmo_rounds += st_host[0];
      lmo_bids += st_host[1];
      solved = st_host[2] == 0;
      if (!solved) ++lmo_fallbacks;
    }
    if (!solved) {
      Mi_host.resize(nm);
      EVREP_CUDA_OK(cudaMemcpyAsync(Mi_host.data(), w.Mi, sizeof(float) * nm, cudaMemcpyDeviceToHost, stream));
      EVREP_CUDA_OK(cudaMemcpyAsync(red_host, w.red, sizeof(double), cudaMemcpyDeviceToHost, stream));
      EVREP_CUDA_OK(cudaStreamSynchronize(stream));
      if (rect) {
        plan_rp.resize((size_t)n + 1);
        plan_col.resize((size_t)w.plan_cap);
        plan_wd.resize((size_t)w.plan_cap);
        plan_wf.resize((size_t)w.plan_cap);
        int nnz = 0;
        const int rcp = transport_plan_host(Mi_host.data(), n, m, w.plan_cap, plan_rp.data(), plan_col.data(), plan_wd.data(), &nnz);
        if (rcp) return rcp;
        for (int e = 0; e < nnz; ++e) plan_wf[(size_t)e] = (float)plan_wd[(size_t)e];
        ++lmo_fallbacks;
        EVREP_CUDA_OK(cudaMemcpyAsync(w.plan_rp, plan_rp.data(), sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice, stream));
        EVREP_CUDA_OK(cudaMemcpyAsync(w.plan_col, plan_col.data(), sizeof(int) * (size_t)std::max(nnz, 1), cudaMemcpyHostToDevice, stream));
        EVREP_CUDA_OK(cudaMemcpyAsync(w.plan_w, plan_wf.data(), sizeof(float) * (size_t)std::max(nnz, 1), cudaMemcpyHostToDevice, stream));
      } else {
        lap_solve(Mi_host.data(), n, sigma);
        EVREP_CUDA_OK(cudaMemcpyAsync(w.sigma, sigma.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, stream));
      }
    }
    if (it == 0) f_val = red_host[0];
    int rc;
    if (rect) {
      // hC1 Gc hC2^T = hC1 X with X^T[j, k] = sum_l Gc[k, l] hC2[j, l] (a vertex has <= n + m - 1 entries): sparse product, pack, contraction
      EVREP_CUDA_OK(cudaMemsetAsync(w.Gc, 0, sizeof(float) * nm, stream));
      k_gwb_plan_scatter<<<(n + 255) / 256, 256, 0, stream>>>(w.plan_rp, w.plan_col, w.plan_w, n, m, w.Gc);
      k_gwb_sparse_xt<<<eb, 256, 0, stream>>>(w.hC2, w.plan_rp, w.plan_col, w.plan_w, m, n, w.Bp);
      rc = launch_gemm_pack(w.Bp, m, n, n, nullptr, w.imgB, stream);
      if (rc) return rc;
      rc = launch_gemm_packed(w.imgA, w.imgB, w.AGc, n, m, n, 1.f, nullptr, nullptr, stream);
    } else {
      // hC1 Gc hC2^T = (1 / n) hC1 (hC2[:, sigma])^T : the gather writes the B operand image, then the tensor-core contraction
      rc = launch_gemm_pack(w.hC2, m, n, m, w.sigma, w.imgB, stream);
      if (rc) return rc;
      rc = launch_gemm_packed(w.imgA, w.imgB, w.AGc, n, m, n, 1.f / (float)n, nullptr, nullptr, stream);
    }
    if (rc) return rc;
    EVREP_CUDA_OK(cudaMemsetAsync(w.red, 0, sizeof(double) * 8, stream));
    k_gwb_linesearch<<<eb, 256, 0, stream>>>(n, m, w.cr, w.cc, w.AG, w.AGc, w.G, w.sigma, rect ? w.Gc : nullptr, w.red);
    EVREP_CUDA_OK(cudaMemcpyAsync(red_host, w.red, sizeof(double) * 5, cudaMemcpyDeviceToHost, stream));
    EVREP_CUDA_OK(cudaStreamSynchronize(stream));
    // f(G + alpha dG) = f(G) + b alpha + a alpha^2 with
    const double a = -red_host[0];
    const double b = red_host[1] - red_host[2] - red_host[3];
    if (!std::isfinite(a) || !std::isfinite(b)) break;
    double alpha;
    if (a > 0) alpha = std::min(1.0, std::max(0.0, -b / (2 * a)));
    else alpha = (a + b < 0) ? 1.0 : 0.0;
    const double old = f_val;
    f_val = old + a * alpha * alpha + b * alpha;
    {  // the next LMO's matrix differs from this one by alpha * (AGc - AG): prices stay 2 max |change| - optimal
      const double adg_max = red_host[4];  // written as the bit pattern of a non-negative double by atomicMax
      warm_eps = (std::isfinite(adg_max) && alpha > 0.0) ? 2.0 * alpha * adg_max : 0.0;
      if (alpha == 0.0) warm_eps = 1e-300;  // same matrix again: the old prices are already optimal
    }
    if (alpha != 0.0) k_gwb_step<<<eb, 256, 0, stream>>>(n, m, (float)alpha, w.sigma, rect ? w.Gc : nullptr, w.AGc, w.G, w.AG);
    EVREP_CUDA_OK(cudaGetLastError());
    // POT's stopping rule (|df| < tol_abs or |df| / |f| < tol_rel), with both tolerances floored at the resolution of
    // the float32 gradient (4 ulp of the loss): below it the predicted decrease is rounding noise and the iteration
    // would wander until max_iter (seen at n = 1000: 10000 steps of ~1e-8 relative "progress")
    const double dlt = fabs(f_val - old);
    const double floor_rel = 4.0 * 1.1920929e-7;
    if (dlt < std::max(tol_abs, floor_rel * fabs(f_val)) || dlt / std::max(fabs(f_val), 1e-300) < std::max(tol_rel, floor_rel)) { ++it; break; }
  }
  // the loss at the final plan, from a fresh contraction (two GEMMs: X^T = hC2 G^T, then hC1 X)
  {
    int rc = launch_gemm_nt_3xtf32(w.hC2, w.G, w.Bp, m, n, m, 1.f, nullptr, nullptr, stream);  // Bp[j, i] = sum_l hC2[j, l] G[i, l]
    if (rc) return rc;
    rc = launch_gemm_nt_3xtf32(w.hC1, w.Bp, w.AG, n, m, n, 1.f, nullptr, nullptr, stream);     // AG[i, j] = sum_k hC1[i, k] Bp[j, k]
    if (rc) return rc;
    EVREP_CUDA_OK(cudaMemsetAsync(w.red, 0, sizeof(double) * 8, stream));
    k_gwb_grad<<<eb, 256, 0, stream>>>(n, m, w.cr, w.cc, w.AG, w.G, w.Mi, w.red);
    EVREP_CUDA_OK(cudaMemcpyAsync(red_host, w.red, sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (T_out) EVREP_CUDA_OK(cudaMemcpyAsync(T_out, w.G, sizeof(float) * nm, cudaMemcpyDeviceToDevice, stream));
    EVREP_CUDA_OK(cudaStreamSynchronize(stream));
  }
  if (gw_dist_host) *gw_dist_host = red_host[0];
  if (iters_host) *iters_host = it;
  if (lmo_stats_host) {
    lmo_stats_host[0] = (int)std::min<long long>(lmo_rounds, INT_MAX);
    lmo_stats_host[1] = (int)std::min<long long>(lmo_bids, INT_MAX);
    lmo_stats_host[2] = (int)lmo_fallbacks;
  }
  return EVREP_OK;
}

}  // namespace evrep
