// Batched bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, fp32 accumulators in TMEM), operands
// staged by TMA (cp.async.bulk.tensor, 128B swizzle) through an mbarrier ring.  gp_bgemm_bf16.
//
//   C[b] (fp32, optional bf16 copy) = alpha * op(A[b]) . op(B[b]) (+ beta * C[b])
//
// Used for the dense (A, X, S, W) contractions of the DiffPool path at sizes where a 128-row MMA
// tile is not mostly padding (encoders.py:319,322,1278,1279,1311 and their backward products).
// Both operands may be K-major or MN-major, so every product on the path reads its operands in
// their natural row-major HBM layout (A.X: X is N-major; A^T.dU: A is M-major; S^T.Z: both
// MN-major; S.S^T: both K-major ...) -- no transposed copies are ever written.
//
// CTA = 6 warps: warps 0-3 epilogue (TMEM lane quarter = warp id), warp 4 TMA producer, warp 5
// TMEM allocator + single-thread MMA issuer.  Tile 128 x BN x 64 (BN = 64/128/256), UMMA 128xBNx16,
// cta_group::1.  One output tile per CTA; two CTAs are co-resident per SM for BN <= 128 so that one
// tile's epilogue overlaps the other's main loop.  Per-graph limits truncate the K loop (tile
// skipping beyond a graph's node count) and zero-fill clipped rows/columns in the epilogue.
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace gp {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 bf16 = 128 B = one swizzle atom row

struct TcParams {
  float* C; __nv_bfloat16* Cb;
  int M, N, K, batch;
  long long ldC, sCb, ldCb, sCbb;
  const int32_t* lim; int lim_m, lim_n, lim_k;
  float alpha, beta; const float* alpha_dev;
  const float* bias; int relu;
  int split_k;
  // EPI == 1 (fused link-prediction loss): bf16 adjacency, bf16 gsym output (Cb), per-warp partial sums
  const __nv_bfloat16* adjb; long long ldadj, sadjb;
  float* partial;
};

constexpr float kEpsLinkTc = 1e-7f;

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (CUDA error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    if (++spins > (1u << 26)) { __trap(); }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, 128B swizzle (cute::UMMA::SmemDescriptor): start>>4 [0,14),
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN, int STAGES>
struct TcSmem {
  static constexpr int kA = BM * BK * 2;        // 16 KB
  static constexpr int kB = BN * BK * 2;
  static constexpr int kStage = kA + kB;
  static constexpr int kBytes = STAGES * kStage + 1024 /*align slack*/ + 256 /*barriers*/;
};

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(192, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  using L = TcSmem<BN, STAGES>;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;        // 1024B aligned (swizzle atoms)
  const uint32_t bar_base = base + STAGES * L::kStage;
  // barriers: full[STAGES], empty[STAGES], accum (8 B each), then the TMEM base address word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * L::kStage + 8 * (2 * STAGES + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = p.split_k > 1 ? p.split_k : 1;
  const int b = blockIdx.z / split, ks = blockIdx.z % split;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  int Me = p.M, Ne = p.N, Ke = p.K;
  if (p.lim != nullptr) {
    const int l = p.lim[b];
    if (p.lim_m) Me = min(Me, l);
    if (p.lim_n) Ne = min(Ne, l);
    if (p.lim_k) Ke = min(Ke, l);
  }
  const bool live = (m0 < Me) && (n0 < Ne) && (Ke > 0);
  const int ktiles = live ? (Ke + BK - 1) / BK : 0;
  const int per = (ktiles + split - 1) / split;
  const int kt0 = min(ktiles, ks * per), kt1 = min(ktiles, kt0 + per);
  const int nk = kt1 - kt0;

  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  if (warp == 5) {
    // TMEM allocation (whole warp), barrier init (one lane)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      mbar_init(accum_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  }
  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 4) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int i = 0; i < nk; ++i) {
        const int s = i % STAGES, ph = (i / STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_expect_tx(full_bar(s), L::kStage);
        const uint32_t sa = base + s * L::kStage, sb = sa + L::kA;
        const int k0 = (kt0 + i) * BK;
        if (A_MN) {     // A stored [K rows, M cols]: two 64-wide M chunks of 64 k-rows each
          tma_load_3d(sa, &tmA, full_bar(s), m0, k0, b);
          tma_load_3d(sa + 8192, &tmA, full_bar(s), m0 + 64, k0, b);
        } else {        // A stored [M rows, K cols]: one 128-row box
          tma_load_3d(sa, &tmA, full_bar(s), k0, m0, b);
        }
        if (B_MN) {     // B stored [K rows, N cols]
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_3d(sb + j * 8192, &tmB, full_bar(s), n0 + 64 * j, k0, b);
        } else {        // B stored [N rows, K cols]
          tma_load_3d(sb, &tmB, full_bar(s), k0, n0, b);
        }
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=BF16, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                             ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int i = 0; i < nk; ++i) {
        const int s = i % STAGES, ph = (i / STAGES) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = base + s * L::kStage, sb = sa + L::kA;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // K-major: +32 B per UMMA_K inside the 128 B swizzled row, SBO = 1024 (8 rows x 128 B)
          // MN-major: +16 k-rows x 128 B = 2048 B per UMMA_K, LBO = 8192 (next 64-wide MN chunk), SBO = 1024
          const uint64_t ad = A_MN ? umma_desc(sa + k * 2048, 8192, 1024) : umma_desc(sa + k * 32, 16, 1024);
          const uint64_t bd = B_MN ? umma_desc(sb + k * 2048, 8192, 1024) : umma_desc(sb + k * 32, 16, 1024);
          tc_mma_bf16(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(empty_bar(s));                 // frees the smem stage when these MMAs retire
      }
      if (nk > 0) tc_commit(accum_bar);          // accumulator complete
    }
  } else {
    // ===== epilogue warps 0..3: TMEM -> registers -> global =====
    const int row = m0 + warp * 32 + lane;
    float alpha = p.alpha;
    if (p.alpha_dev != nullptr) alpha *= *p.alpha_dev;
    if (nk > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
    const bool row_ok = row < p.M;
    const bool row_in = row < Me;
    if constexpr (EPI == 1) {
      // ---- fused link-prediction loss (encoders.py:1311-1331): the accumulator tile is P = S S^T.
      //   l = -a log(p+eps) - (1-a) log(1-p+eps) summed over the nb x nb block;
      //   G = dl/dp evaluated with a[m,n] only (bf16): the backward is dS = (G + G^T).S, computed as two
      //   GEMMs G.S + G^T.S (G^T is just the M-major view of G), so the transposed adjacency is never read.
      float lsum = 0.f;
      const __nv_bfloat16* ab = p.adjb + (long long)b * p.sadjb;
      __nv_bfloat16* grow = p.Cb != nullptr ? p.Cb + (long long)b * p.sCbb + (long long)row * p.ldCb : nullptr;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        if (nk > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        const int nbase = n0 + c * 32;
        if (nk == 0 || !row_ok || nbase >= p.N) continue;
        // this thread's 32 adjacency entries a[row, nbase..nbase+31] (64 contiguous bytes)
        __align__(16) __nv_bfloat16 a[32];
        if (row_in && nbase + 32 <= p.N && (p.ldadj % 8 == 0)) {
          const uint4* src = reinterpret_cast<const uint4*>(ab + (long long)row * p.ldadj + nbase);
#pragma unroll
          for (int j = 0; j < 4; ++j) reinterpret_cast<uint4*>(a)[j] = src[j];
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            a[j] = (row_in && nbase + j < p.N) ? ab[(long long)row * p.ldadj + nbase + j] : __float2bfloat16_rn(0.f);
        }
        float g[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = nbase + j;
          float gg = 0.f;
          if (row_in && n < Ne) {
            float pv = __uint_as_float(v[j]);
            const bool over = pv > 1.f;
            if (over) pv = 1.f;
            const float a1 = __bfloat162float(a[j]);
            const float pe = pv + kEpsLinkTc, qe = 1.f - pv + kEpsLinkTc;
            lsum -= a1 * __logf(pe) + (1.f - a1) * __logf(qe);
            if (!over) gg = -a1 * __fdividef(1.f, pe) + (1.f - a1) * __fdividef(1.f, qe);
          }
          g[j] = gg;
        }
        if (grow != nullptr) {
          if (nbase + 32 <= p.N && (p.ldCb % 8 == 0)) {
            uint4* dst = reinterpret_cast<uint4*>(grow + nbase);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __nv_bfloat162 h0 = __floats2bfloat162_rn(g[8 * j], g[8 * j + 1]);
              __nv_bfloat162 h1 = __floats2bfloat162_rn(g[8 * j + 2], g[8 * j + 3]);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(g[8 * j + 4], g[8 * j + 5]);
              __nv_bfloat162 h3 = __floats2bfloat162_rn(g[8 * j + 6], g[8 * j + 7]);
              uint4 o;
              o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
              o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
              dst[j] = o;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nbase + j < p.N) grow[nbase + j] = __float2bfloat16_rn(g[j]);
          }
        }
      }
      lsum = warp_sum(lsum);
      if (lane == 0)
        p.partial[(((long long)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 4 + warp] = lsum;
    } else {
    float* crow = p.C != nullptr ? p.C + (long long)b * p.sCb + (long long)row * p.ldC : nullptr;
    __nv_bfloat16* cbrow = p.Cb != nullptr ? p.Cb + (long long)b * p.sCbb + (long long)row * p.ldCb : nullptr;
    const bool vec4 = (p.ldC % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (p.sCb % 4 == 0);
    const bool vec8 = (p.ldCb % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.Cb) & 15) == 0) && (p.sCbb % 8 == 0);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      if (nk > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      const int nbase = n0 + c * 32;
      if (!row_ok || nbase >= p.N) continue;
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int n = nbase + j;
        float x = (row_in && n < Ne) ? alpha * __uint_as_float(v[j]) : 0.f;
        f[j] = x;
      }
      if (split > 1) {
        if (row_in) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nbase + j < Ne && f[j] != 0.f) atomicAdd(crow + nbase + j, f[j]);
        }
        continue;
      }
      if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (nbase + j < p.N) f[j] += p.bias[nbase + j];
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
      }
      const bool full = nbase + 32 <= p.N;
      if (crow != nullptr) {
        if (full && vec4) {
          float4* dst = reinterpret_cast<float4*>(crow + nbase);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            if (p.beta != 0.f) {
              const float4 old = dst[j];
              o.x += p.beta * old.x; o.y += p.beta * old.y; o.z += p.beta * old.z; o.w += p.beta * old.w;
              f[4 * j] = o.x; f[4 * j + 1] = o.y; f[4 * j + 2] = o.z; f[4 * j + 3] = o.w;
            }
            dst[j] = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (nbase + j < p.N) {
              if (p.beta != 0.f) f[j] += p.beta * crow[nbase + j];
              crow[nbase + j] = f[j];
            }
          }
        }
      }
      if (cbrow != nullptr) {
        if (full && vec8) {
          uint4* dst = reinterpret_cast<uint4*>(cbrow + nbase);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(f[8 * j], f[8 * j + 1]);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
            o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
            dst[j] = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nbase + j < p.N) cbrow[nbase + j] = __float2bfloat16_rn(f[j]);
        }
      }
    }
    }  // EPI == 0
  }

  // teardown: everyone done with TMEM before the allocating warp frees it
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 -> bf16 conversion (strided rows, optional zero padding of the row tail up to ld_out)
// ------------------------------------------------------------------------------------------------
__global__ void cvt_bf16_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y,
                                long long ldy, long long rows, int cols, int cols_pad) {
  const int cpr = (cols_pad + 7) / 8;                       // 8-element chunks per row
  const long long total = rows * cpr;
  const bool vec = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (ldy % 8 == 0) &&
                   ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cpr;
    const int c0 = (int)(i - r * cpr) * 8;
    const float* src = x + r * ldx + c0;
    __nv_bfloat16* dst = y + r * ldy + c0;
    if (vec && c0 + 8 <= cols) {
      const float4 a = *reinterpret_cast<const float4*>(src);
      const float4 c = *reinterpret_cast<const float4*>(src + 4);
      __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(c.x, c.y), h3 = __floats2bfloat162_rn(c.z, c.w);
      uint4 o;
      o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
      o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(dst) = o;
    } else {
      for (int j = 0; j < 8 && c0 + j < cols_pad; ++j)
        dst[j] = __float2bfloat16_rn(c0 + j < cols ? src[j] : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 3-D bf16 tensor [batch][rows][cols] (cols contiguous), box = {64, box_rows, 1}, 128B swizzle
static int make_map(CUtensorMap* tm, const void* ptr, long long cols, long long rows, long long batch,
                    long long ld, long long sb, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) return fail(GP_ERR_CUDA, "cuTensorMapEncodeTiled not available");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? sb : ld * rows) * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(GP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): cols=%lld rows=%lld batch=%lld ld=%lld sb=%lld",
                                     (int)r, cols, rows, batch, ld, sb);
  return GP_OK;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI = 0>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, cudaStream_t st) {
  using L = TcSmem<BN, STAGES>;
  auto kern = tc_gemm_kernel<BN, STAGES, A_MN, B_MN, EPI>;
  static bool configured = false;
  if (!configured) {
    GP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes));
    configured = true;
  }
  const int split = p.split_k > 1 ? p.split_k : 1;
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, p.batch * split);
  kern<<<grid, 192, L::kBytes, st>>>(tmA, tmB, p);
  GP_LAUNCHED();
  return GP_OK;
}

template <int BN, int STAGES>
static int dispatch_major(int a_mn, int b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p,
                          cudaStream_t st) {
  if (!a_mn && !b_mn) return launch_tc<BN, STAGES, false, false>(tmA, tmB, p, st);
  if (!a_mn && b_mn) return launch_tc<BN, STAGES, false, true>(tmA, tmB, p, st);
  if (a_mn && !b_mn) return launch_tc<BN, STAGES, true, false>(tmA, tmB, p, st);
  return launch_tc<BN, STAGES, true, true>(tmA, tmB, p, st);
}

__global__ void scale_fill_kernel2(float* c, long long sCb, long long ldC, int M, int N, int batch, float beta) {
  const long long total = (long long)batch * M * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const long long r = i / N;
    const int m = (int)(r % M);
    const long long bb = r / M;
    float* q = c + bb * sCb + (long long)m * ldC + n;
    *q = beta == 0.f ? 0.f : beta * (*q);
  }
}

}  // namespace gp

using namespace gp;

extern "C" int gp_bgemm_bf16(const gp_gemm_bf16* g, gp_stream_t stream) {
  GP_REQUIRE(g != nullptr, "bgemm_bf16: null descriptor");
  GP_REQUIRE(g->A && g->B && (g->C || g->Cb), "bgemm_bf16: null operand");
  GP_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0 && g->batch > 0, "bgemm_bf16: bad dims");
  GP_REQUIRE(g->ldA % 8 == 0 && g->ldB % 8 == 0 && (g->batch == 1 || (g->sAb % 8 == 0 && g->sBb % 8 == 0)),
             "bgemm_bf16: operand strides must be multiples of 8 elements (TMA 16-byte rule)");
  GP_REQUIRE((reinterpret_cast<uintptr_t>(g->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g->B) & 15) == 0,
             "bgemm_bf16: operand base must be 16-byte aligned");
  const int split = g->split_k > 1 ? g->split_k : 1;
  GP_REQUIRE((long long)g->batch * split <= 65535, "bgemm_bf16: batch*split_k > 65535");
  GP_REQUIRE(split == 1 || (g->C && !g->Cb && !g->bias && !g->relu), "bgemm_bf16: split_k needs fp32 C only");
  cudaStream_t st = S(stream);

  CUtensorMap tmA, tmB;
  const int BN = g->N > 128 ? 256 : (g->N > 64 ? 128 : 64);
  if (g->a_major == 0) GP_TRY(make_map(&tmA, g->A, g->K, g->M, g->batch, g->ldA, g->sAb, BM));
  else                 GP_TRY(make_map(&tmA, g->A, g->M, g->K, g->batch, g->ldA, g->sAb, BK));
  if (g->b_major == 0) GP_TRY(make_map(&tmB, g->B, g->K, g->N, g->batch, g->ldB, g->sBb, BN));
  else                 GP_TRY(make_map(&tmB, g->B, g->N, g->K, g->batch, g->ldB, g->sBb, BK));

  TcParams p;
  p.C = g->C; p.Cb = reinterpret_cast<__nv_bfloat16*>(g->Cb);
  p.M = g->M; p.N = g->N; p.K = g->K; p.batch = g->batch;
  p.ldC = g->ldC; p.sCb = g->sCb; p.ldCb = g->ldCb; p.sCbb = g->sCbb;
  p.lim = g->lim; p.lim_m = g->lim_m; p.lim_n = g->lim_n; p.lim_k = g->lim_k;
  p.alpha = g->alpha; p.beta = g->beta; p.alpha_dev = g->alpha_dev;
  p.bias = g->bias; p.relu = g->relu; p.split_k = g->split_k;
  p.adjb = nullptr; p.ldadj = p.sadjb = 0; p.partial = nullptr;
  if (split > 1 && p.beta != 1.f) {
    const long long total = (long long)g->batch * g->M * g->N;
    int blocks = (int)((total + 255) / 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    scale_fill_kernel2<<<blocks, 256, 0, st>>>(p.C, p.sCb, p.ldC, p.M, p.N, p.batch, p.beta);
    GP_LAUNCHED();
    p.beta = 1.f;
  }
  if (BN == 256) return dispatch_major<256, 4>(g->a_major, g->b_major, tmA, tmB, p, st);
  if (BN == 128) return dispatch_major<128, 3>(g->a_major, g->b_major, tmA, tmB, p, st);
  return dispatch_major<64, 4>(g->a_major, g->b_major, tmA, tmB, p, st);
}

extern "C" int gp_cvt_f32_bf16(const float* x, long long ldx, void* y, long long ldy, long long rows, int cols,
                               int cols_pad, gp_stream_t stream) {
  GP_REQUIRE(x && y && rows > 0 && cols > 0 && cols_pad >= cols && ldy >= cols_pad && ldx >= cols, "cvt_bf16: bad args");
  const long long total = rows * ((cols_pad + 7) / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  cvt_bf16_kernel<<<(int)blocks, 256, 0, S(stream)>>>(x, ldx, reinterpret_cast<__nv_bfloat16*>(y), ldy, rows, cols,
                                                      cols_pad);
  GP_LAUNCHED();
  return GP_OK;
}

// Fused link-prediction loss on tensor cores: P = S S^T tiles (128 x 256) live only in TMEM; the
// epilogue does the masked BCE against the bf16 adjacency, writes gsym (bf16) and one partial sum
// per epilogue warp.  n_partial = batch * ceil(N/128) * ceil(N/256) * 4.
// (superseded by the persistent v2 kernel in gemm_tc2.cu, which exports gp_linkloss_tc; kept as the
// single-tile-per-CTA reference implementation of the same epilogue)
extern "C" int gp_linkloss_tc_v1(const void* s_bf16, long long lds, const void* adj_bf16, long long ldadj,
                                 const int32_t* nb, int B, int N, int K, float* partial, void* gsym_bf16,
                                 long long ldg, gp_stream_t stream) {
  GP_REQUIRE(s_bf16 && adj_bf16 && partial && B > 0 && N > 0 && K > 0, "linkloss_tc: bad args");
  GP_REQUIRE(lds % 8 == 0 && ldadj % 8 == 0 && (gsym_bf16 == nullptr || ldg >= N), "linkloss_tc: bad strides");
  GP_REQUIRE(B <= 65535, "linkloss_tc: batch too large");
  CUtensorMap tmA, tmB;
  GP_TRY(make_map(&tmA, s_bf16, K, N, B, lds, (long long)N * lds, BM));
  GP_TRY(make_map(&tmB, s_bf16, K, N, B, lds, (long long)N * lds, 256));
  TcParams p;
  p.C = nullptr; p.Cb = reinterpret_cast<__nv_bfloat16*>(gsym_bf16);
  p.M = N; p.N = N; p.K = K; p.batch = B;
  p.ldC = p.sCb = 0; p.ldCb = ldg; p.sCbb = (long long)N * ldg;
  p.lim = nb; p.lim_m = p.lim_n = nb != nullptr; p.lim_k = 0;
  p.alpha = 1.f; p.beta = 0.f; p.alpha_dev = nullptr; p.bias = nullptr; p.relu = 0; p.split_k = 0;
  p.adjb = reinterpret_cast<const __nv_bfloat16*>(adj_bf16); p.ldadj = ldadj; p.sadjb = (long long)N * ldadj;
  p.partial = partial;
  return launch_tc<256, 4, false, false, 1>(tmA, tmB, p, S(stream));
}
