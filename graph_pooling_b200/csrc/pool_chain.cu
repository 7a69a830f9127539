// Chained pooling contraction on tensor cores (encoders.py:1279):   A'[b] = S[b]^T . A[b] . S[b]   in ONE launch.
//
// The intermediate T = S^T A never makes a round trip through HBM: a T tile is accumulated in TMEM (first GEMM),
// converted to bf16 into SHARED memory in the swizzled K-major layout a UMMA A-operand descriptor expects, and is the
// A operand of a second tcgen05.mma that accumulates the A' row block in another TMEM region.
//
//   work item      = (graph b, block of 128 clusters m0 .. m0+127)                  -> rows m0.. of T and of A'
//   first GEMM     : T_j [128 x 128]  = sum_k S[k, m0..]^T . A[k, 128 j ..]          (k over the n_b real nodes)
//   second GEMM    : A'[128 x Kc]    += T_j [128 x 128] . S[128 j .., c0 .. c0+Kc]
//
// TMEM budget (512 columns per SM): A' accumulator Kc <= 256 columns + two T accumulators of 128 columns (the
// epilogue of T_j overlaps the MMAs of T_j+1).  K <= 256 clusters: one CTA owns all columns of its A' row block.
// 256 < K <= 512 (cfg4: K = 512): a CLUSTER OF TWO CTAs shares a work item -- CTA c accumulates the A' columns
// [c*Kc, (c+1)*Kc) and computes the T tiles j with j % 2 == c; every bf16 T tile is pushed into the partner's shared
// memory with one bulk DSMEM copy (cp.async.bulk.shared::cluster, completing on the partner's mbarrier), so each T
// tile is computed ONCE per cluster and both CTAs multiply it with their own column half of S.
//
// Training also needs T in the backward (dS += T^T dA'): the same shared-memory tile is written out once as bf16 by a
// TMA store (cp.async.bulk.tensor ... global.shared::cta); inference passes t = NULL and T never exists in HBM.
//
// Warps: 0-3 epilogue (TMEM -> bf16 shared tile; final A' block -> global), 4 TMA producer, 5 MMA issuer,
// 6 exchange (T store, DSMEM push, "slot free" relay to the partner).  Operands keep their natural row-major layout
// (S [B,N,K], A [B,N,N]); all three operand roles use 64 x 64 TMA boxes with the 128-byte swizzle.
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>
#include "common.cuh"

namespace gp {
namespace chain {

constexpr int BM = 128;          // clusters per work item (rows of T and A')
constexpr int BN1 = 128;         // node columns of one T tile
constexpr int BK = 64;           // contraction step
constexpr int STAGES = 4;
constexpr int kStage = 32768;    // first GEMM: 16 KB S^T tile + 16 KB A tile; second GEMM: up to 64 x 256 of S
constexpr int kSlot = 32768;     // one bf16 T tile, two 64-column halves of 16 KB
constexpr int kEpiWarps = 4;
constexpr int kThreads = (kEpiWarps + 3) * 32;
constexpr int kSmemBytes = STAGES * kStage + 2 * kSlot + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kTmemCols = 512;
constexpr int kAccT = 256;       // TMEM column of the first T accumulator (A' occupies 0..255)

struct Maps { CUtensorMap s, adj, t; };

struct Params {
  int B, N, K;
  const int32_t* nb; const int32_t* order;
  int tiles_m; long long total_work; int nclusters;
  int ng2;                       // columns of A' per CTA (multiple of 64, <= 256)
  int store_t;
  float* ap; long long ldap; __nv_bfloat16* apb; long long ldapb;
};

struct Work { int b, m0, Ne, NT, KT, R; };

// ---- PTX wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
  return ok != 0;
}
// waits that may observe a barrier completed by the partner CTA acquire at cluster scope
__device__ __forceinline__ bool mbar_try_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    if (++spins > (1u << 22)) { __trap(); }     // protocol bug: fail loudly instead of hanging the GPU
  }
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_cluster(bar, parity)) {
    if (++spins > (1u << 22)) { __trap(); }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// shared::cta -> partner's shared memory, completion (bytes) on the PARTNER's mbarrier
__device__ __forceinline__ void dsmem_push(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
// UMMA shared-memory descriptor, SWIZZLE_128B (same encoding as gemm_tc2.cu)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}

__device__ __forceinline__ Work get_work(const Params& p, long long w) {
  Work k;
  int bi = (int)(w / p.tiles_m);
  const int mt = (int)(w - (long long)bi * p.tiles_m);
  if (p.order != nullptr) bi = p.order[bi];
  k.b = bi; k.m0 = mt * BM;
  k.Ne = p.nb != nullptr ? max(0, min(p.N, p.nb[bi])) : p.N;
  k.NT = (k.Ne + BN1 - 1) / BN1;
  k.KT = (k.Ne + BK - 1) / BK;
  return k;
}

// CL = CTAs per work item (1: K <= 256, 2: K <= 512)
template <int CL>
__global__ void __launch_bounds__(kThreads, 1)
pool_chain_kernel(const __grid_constant__ Maps maps, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t slot_base = base + STAGES * kStage;
  const uint32_t bar_base = slot_base + 2 * kSlot;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };       // T accumulator a is complete
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };  // ... has been read by the epilogue
  auto sfull_bar = [&](int c) { return bar_base + 8u * (2 * STAGES + 4 + c); };   // slot c holds the bf16 T tile of CTA c
  auto sfree_bar = [&](int c) { return bar_base + 8u * (2 * STAGES + 6 + c); };   // this CTA's second GEMM has read slot c
  const uint32_t xfree_bar = bar_base + 8u * (2 * STAGES + 8);    // the PARTNER has consumed the tile pushed to it
  const uint32_t stored_bar = bar_base + 8u * (2 * STAGES + 9);   // exchange warp is done reading the own slot
  const uint32_t afull_bar = bar_base + 8u * (2 * STAGES + 10);   // A' block complete
  const uint32_t aempty_bar = bar_base + 8u * (2 * STAGES + 11);  // A' block drained
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 12);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * kStage + 2 * kSlot + 8 * (2 * STAGES + 12));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kProd = kEpiWarps, kMma = kEpiWarps + 1, kXch = kEpiWarps + 2;
  uint32_t rank = 0;
  if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const long long cluster_id = blockIdx.x / CL;

  if (warp == kMma) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kEpiWarps); }
      for (int c = 0; c < 2; ++c) {
        mbar_init(sfull_bar(c), c == (int)rank ? kEpiWarps : 1);   // own: epilogue warps; partner's: expect_tx arrival
        mbar_init(sfree_bar(c), 1);
      }
      mbar_init(xfree_bar, 1); mbar_init(stored_bar, 1); mbar_init(afull_bar, 1); mbar_init(aempty_bar, kEpiWarps);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      fence_async_smem();
    }
  }
  if (warp == kProd && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.s)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.adj)) : "memory");
    if (p.store_t) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.t)) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                       // the partner's barriers exist before anything remote happens
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  const int ng2 = p.ng2, col0 = (int)rank * ng2;

  if (warp == kProd) {
    // ===== TMA producer: the stage sequence below is mirrored exactly by the MMA issuer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (long long w = cluster_id; w < p.total_work; w += p.nclusters) {
        Work k = get_work(p, w);
        const int R = (k.NT + CL - 1) / CL;
        for (int r = 0; r <= R; ++r) {
          const int j = r * CL + (int)rank;
          if (r < R && j < k.NT) {
            for (int g = 0; g < k.KT; ++g, ++it) {
              const int s = it % STAGES, ph = (it / STAGES) & 1;
              mbar_wait(empty_bar(s), ph ^ 1);
              mbar_expect_tx(full_bar(s), kStage);
              const uint32_t sa = base + s * kStage, sb = sa + 16384;
              tma_load_3d(sa, &maps.s, full_bar(s), k.m0, g * BK, k.b);              // S[k rows, m0.. cols] = (S^T) tile
              tma_load_3d(sa + 8192, &maps.s, full_bar(s), k.m0 + 64, g * BK, k.b);
              tma_load_3d(sb, &maps.adj, full_bar(s), j * BN1, g * BK, k.b);         // A[k rows, node cols]
              tma_load_3d(sb + 8192, &maps.adj, full_bar(s), j * BN1 + 64, g * BK, k.b);
            }
          }
          if (r >= 1) {
            for (int c = 0; c < CL; ++c) {
              const int t = (r - 1) * CL + c;
              if (t >= k.NT) continue;
              const int nh = (min(BN1, k.Ne - t * BN1) + BK - 1) / BK;
              for (int h = 0; h < nh; ++h, ++it) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(empty_bar(s), ph ^ 1);
                mbar_expect_tx(full_bar(s), (uint32_t)ng2 * 128u);
                const uint32_t sa = base + s * kStage;
                for (int jj = 0; jj < ng2 / 64; ++jj)                                // S[node rows, col0.. cols]
                  tma_load_3d(sa + jj * 8192, &maps.s, full_bar(s), col0 + 64 * jj, t * BN1 + h * BK, k.b);
              }
            }
          }
        }
      }
    }
  } else if (warp == kMma) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t it = 0, tc = 0, ac = 0;
      uint32_t fills[2] = {0u, 0u};
      const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                              ((uint32_t)(BN1 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);        // A, B both MN-major
      const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) |
                              ((uint32_t)(ng2 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);        // A K-major, B MN-major
      for (long long w = cluster_id; w < p.total_work; w += p.nclusters) {
        Work k = get_work(p, w);
        const int R = (k.NT + CL - 1) / CL;
        bool first_g2 = true;
        for (int r = 0; r <= R; ++r) {
          const int j = r * CL + (int)rank;
          if (r < R && j < k.NT) {
            const uint32_t a = tc & 1, aph = (tc >> 1) & 1;
            mbar_wait(tempty_bar(a), aph ^ 1);
            tc_fence_after();
            const uint32_t tm = tmem_base + kAccT + a * BN1;
            for (int g = 0; g < k.KT; ++g, ++it) {
              const int s = it % STAGES, ph = (it / STAGES) & 1;
              mbar_wait(full_bar(s), ph);
              tc_fence_after();
              const uint32_t sa = base + s * kStage, sb = sa + 16384;
#pragma unroll
              for (int kk = 0; kk < BK / 16; ++kk)
                tc_mma_bf16(tm, umma_desc(sa + kk * 2048, 8192, 1024), umma_desc(sb + kk * 2048, 8192, 1024), idesc1,
                            (g > 0 || kk > 0) ? 1u : 0u);
              tc_commit(empty_bar(s));
            }
            tc_commit(tfull_bar(a));
            ++tc;
          }
          if (r >= 1) {
            for (int c = 0; c < CL; ++c) {
              const int t = (r - 1) * CL + c;
              if (t >= k.NT) continue;
              if (first_g2) {                              // the previous work item's A' block has been drained
                mbar_wait(aempty_bar, (ac & 1) ^ 1);
                tc_fence_after();
              }
              if (c != (int)rank) {
                mbar_expect_tx(sfull_bar(c), kSlot);       // arms this phase; the partner's push may already have landed
                mbar_wait_cluster(sfull_bar(c), fills[c] & 1);
              } else {
                mbar_wait(sfull_bar(c), fills[c] & 1);
              }
              tc_fence_after();
              const uint32_t slot = slot_base + c * kSlot;
              const int nh = (min(BN1, k.Ne - t * BN1) + BK - 1) / BK;
              for (int h = 0; h < nh; ++h, ++it) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t sa = base + s * kStage;
#pragma unroll
                for (int kk = 0; kk < BK / 16; ++kk)
                  tc_mma_bf16(tmem_base, umma_desc(slot + h * 16384 + kk * 32, 16, 1024),
                              umma_desc(sa + kk * 2048, 8192, 1024), idesc2, (first_g2 && h == 0 && kk == 0) ? 0u : 1u);
                tc_commit(empty_bar(s));
              }
              first_g2 = false;
              tc_commit(sfree_bar(c));
              ++fills[c];
            }
          }
        }
        if (k.NT > 0) { tc_commit(afull_bar); ++ac; }
      }
    }
  } else if (warp == kXch) {
    // ===== exchange warp: T tile -> HBM (training), T tile -> partner, "your tile has been consumed" relay =====
    if (lane == 0) {
      uint32_t ownfills = 0, peerfills = 0;
      const uint32_t own_slot = slot_base + rank * kSlot;
      uint32_t peer_slot = 0, peer_sfull = 0, peer_xfree = 0;
      if (CL > 1) {
        peer_slot = map_to_rank(own_slot, rank ^ 1u);
        peer_sfull = map_to_rank(sfull_bar((int)rank), rank ^ 1u);
        peer_xfree = map_to_rank(xfree_bar, rank ^ 1u);
      }
      for (long long w = cluster_id; w < p.total_work; w += p.nclusters) {
        Work k = get_work(p, w);
        const int R = (k.NT + CL - 1) / CL;
        for (int r = 0; r < R; ++r) {
          const int j = r * CL + (int)rank;
          if (j < k.NT) {
            mbar_wait(sfull_bar((int)rank), ownfills & 1);
            if (p.store_t) {
              tma_store_3d(&maps.t, own_slot, j * BN1, k.m0, k.b);
              if (j * BN1 + 64 < p.N) tma_store_3d(&maps.t, own_slot + 16384, j * BN1 + 64, k.m0, k.b);
              bulk_commit();
            }
            if (CL > 1) {
              if (ownfills >= 1) mbar_wait_cluster(xfree_bar, (ownfills - 1) & 1);   // partner's copy of the previous tile is consumed
              dsmem_push(peer_slot, own_slot, kSlot, peer_sfull);
            }
            if (p.store_t) bulk_wait_read();
            mbar_arrive(stored_bar);
            ++ownfills;
          }
          if (CL > 1) {
            const int jp = r * CL + (int)(rank ^ 1u);
            if (jp < k.NT) {
              mbar_wait(sfree_bar((int)(rank ^ 1u)), peerfills & 1);   // own second GEMM is done with the partner's tile
              mbar_arrive_remote(peer_xfree);
              ++peerfills;
            }
          }
        }
      }
      if (p.store_t) bulk_wait_all();
    }
  } else {
    // ===== epilogue warps 0..3: warp q owns TMEM lanes 32q .. 32q+31 =====
    const int quarter = warp & 3;
    const int trow = quarter * 32 + lane;                 // row of the tile (cluster m0 + trow)
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const uint32_t own_slot_off = STAGES * kStage + rank * kSlot;
    uint8_t* slot_gen = smem_gen + own_slot_off;
    uint32_t tc = 0, ownfills = 0, ac = 0;
    const bool vecC = p.ap != nullptr && (reinterpret_cast<uintptr_t>(p.ap) & 15) == 0 && (p.ldap & 3) == 0 &&
                      (((long long)p.K * p.ldap) & 3) == 0;
    const bool vecCb = p.apb != nullptr && (reinterpret_cast<uintptr_t>(p.apb) & 15) == 0 && (p.ldapb & 7) == 0 &&
                       (((long long)p.K * p.ldapb) & 7) == 0;
    for (long long w = cluster_id; w < p.total_work; w += p.nclusters) {
      Work k = get_work(p, w);
      const int R = (k.NT + CL - 1) / CL;
      for (int r = 0; r < R; ++r) {
        const int j = r * CL + (int)rank;
        if (j >= k.NT) continue;
        const uint32_t a = tc & 1, aph = (tc >> 1) & 1;
        mbar_wait(tfull_bar(a), aph);
        tc_fence_after();
        uint32_t pk[64];                                   // the row's 128 T values as packed bf16
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint32_t v[32];
          tmem_ld32(tmem_base + lane_sel + (uint32_t)(kAccT + a * BN1 + c4 * 32), v);
#pragma unroll
          for (int e = 0; e < 16; ++e)
            pk[c4 * 16 + e] = pack_bf16x2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(a));         // the tensor cores may refill this accumulator
        ++tc;
        if (ownfills >= 1) {                               // the previous tile in the own slot is no longer needed by ...
          mbar_wait(sfree_bar((int)rank), (ownfills - 1) & 1);                 // ... the own second GEMM
          mbar_wait(stored_bar, (ownfills - 1) & 1);                           // ... the T store / the push's issue
          if (CL > 1) mbar_wait_cluster(xfree_bar, (ownfills - 1) & 1);        // ... the push itself (partner consumed it)
        }
        // K-major SWIZZLE_128B tile [128 rows][64 cols] x 2 halves: 16-byte chunk index XOR (row & 7)
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int n = q * 8, half = n >> 6, ch = (n & 63) >> 3;
          uint4* dst = reinterpret_cast<uint4*>(slot_gen + half * 16384 + trow * 128 + ((ch ^ (trow & 7)) << 4));
          *dst = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
        fence_async_smem();                                // generic-proxy writes -> visible to UMMA / TMA / bulk copy
        __syncwarp();
        if (lane == 0) mbar_arrive(sfull_bar((int)rank));
        ++ownfills;
      }
      // ---- A' row block: rows m0 + trow, columns col0 .. col0 + ng2 ----
      const int row = k.m0 + trow;
      const bool has_acc = k.NT > 0;
      if (has_acc) {
        mbar_wait(afull_bar, ac & 1);
        tc_fence_after();
      }
      float* crow = p.ap != nullptr ? p.ap + ((long long)k.b * p.K + row) * p.ldap : nullptr;
      __nv_bfloat16* cbrow = p.apb != nullptr ? p.apb + ((long long)k.b * p.K + row) * p.ldapb : nullptr;
#pragma unroll 1
      for (int c = 0; c < ng2 / 32; ++c) {
        const int nbase = col0 + c * 32;
        if (nbase >= p.K) break;                           // warp-uniform
        uint32_t v[32];
        if (has_acc) {
          tmem_ld32(tmem_base + lane_sel + (uint32_t)(c * 32), v);
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = 0u;
        }
        if (row >= p.K) continue;
        const bool full = nbase + 32 <= p.K;
        if (crow != nullptr) {
          if (vecC && full) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<uint4*>(crow + nbase + 4 * q) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          } else {
            for (int e = 0; e < 32 && nbase + e < p.K; ++e) crow[nbase + e] = __uint_as_float(v[e]);
          }
        }
        if (cbrow != nullptr) {
          if (vecCb && full) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4*>(cbrow + nbase + 8 * q) =
                  make_uint4(pack_bf16x2(__uint_as_float(v[8 * q]), __uint_as_float(v[8 * q + 1])),
                             pack_bf16x2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3])),
                             pack_bf16x2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5])),
                             pack_bf16x2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7])));
          } else {
            for (int e = 0; e < 32 && nbase + e < p.K; ++e) cbrow[nbase + e] = __float2bfloat16_rn(__uint_as_float(v[e]));
          }
        }
      }
      if (has_acc) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(aempty_bar);
        ++ac;
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                       // nobody exits while the partner may still push / arrive here
  if (warp == kMma) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---- host ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* q = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &st) == cudaSuccess &&
        st == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(q);
  }
  return fn;
}
// bf16 [batch][rows][cols] with row stride ld (elements): boxes of 64 columns x box_rows rows, 128-byte swizzle
static int make_map(CUtensorMap* tm, const void* ptr, long long cols, long long rows, long long batch, long long ld,
                    long long sb, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) return fail(GP_ERR_CUDA, "cuTensorMapEncodeTiled not available");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? sb : ld * rows) * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(GP_ERR_CUDA, "pool_chain: cuTensorMapEncodeTiled failed (%d): cols=%lld rows=%lld batch=%lld ld=%lld",
                (int)r, cols, rows, batch, ld);
  return GP_OK;
}

template <int CL>
static int launch(const Maps& maps, Params& p, cudaStream_t st) {
  auto kern = pool_chain_kernel<CL>;
  GP_CONFIG_ONCE(GP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int ncl = kNumSMs / CL;
  if (CL > 1) {
    // persistent clusters must all be co-resident (a waiting cluster would deadlock nobody, but grid sizing by the
    // real capacity keeps the static round-robin balanced): ask the runtime, once per device
    static int max_clusters[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (max_clusters[dev & 63] == 0) {
      cfg.gridDim = dim3(kNumSMs / CL * CL);
      int mc = 0;
      if (cudaOccupancyMaxActiveClusters(&mc, kern, &cfg) != cudaSuccess || mc <= 0) { mc = kNumSMs / CL / 2; cudaGetLastError(); }
      max_clusters[dev & 63] = mc;
      if (getenv("GP_DEBUG")) fprintf(stderr, "[gp] pool_chain: cluster %d, smem %d B -> max active clusters %d\n", CL, kSmemBytes, mc);
    }
    ncl = max_clusters[dev & 63];
  }
  if ((long long)ncl > p.total_work) ncl = (int)p.total_work;
  p.nclusters = ncl;
  cfg.gridDim = dim3(ncl * CL);
  GP_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, p));
  GP_LAUNCHED();
  return GP_OK;
}

int run(const void* s_bf16, long long lds, const void* adj_bf16, long long ldadj, const int32_t* nb,
        const int32_t* order, int B, int N, int K, void* t_bf16, long long ldt, float* ap, long long ldap,
        void* ap_bf16, long long ldapb, cudaStream_t st) {
  GP_REQUIRE(s_bf16 && adj_bf16 && B > 0 && N > 0 && K > 0, "pool_chain: bad args");
  GP_REQUIRE(K <= 512, "pool_chain: at most 512 clusters (the A' row block must fit the TMEM of one CTA pair)");
  GP_REQUIRE(ap || ap_bf16, "pool_chain: no output");
  GP_REQUIRE(lds % 8 == 0 && lds >= K && ldadj % 8 == 0 && ldadj >= N, "pool_chain: operand row strides must be multiples of 8");
  GP_REQUIRE((reinterpret_cast<uintptr_t>(s_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(adj_bf16) & 15) == 0,
             "pool_chain: operand bases must be 16-byte aligned");
  GP_REQUIRE(t_bf16 == nullptr || (ldt % 8 == 0 && ldt >= N && (reinterpret_cast<uintptr_t>(t_bf16) & 15) == 0),
             "pool_chain: t must be 16-byte aligned with a row stride that is a multiple of 8");
  GP_REQUIRE((ap == nullptr || ldap >= K) && (ap_bf16 == nullptr || ldapb >= K), "pool_chain: output row stride < K");
  Maps maps;
  Params p;
  GP_TRY(make_map(&maps.s, s_bf16, K, N, B, lds, (long long)N * lds, BK));
  GP_TRY(make_map(&maps.adj, adj_bf16, N, N, B, ldadj, (long long)N * ldadj, BK));
  if (t_bf16 != nullptr) GP_TRY(make_map(&maps.t, t_bf16, N, K, B, ldt, (long long)K * ldt, BM));
  else maps.t = maps.s;
  const int CL = K > 256 ? 2 : 1;
  p.B = B; p.N = N; p.K = K; p.nb = nb; p.order = order;
  p.tiles_m = (K + BM - 1) / BM;
  p.total_work = (long long)p.tiles_m * B;
  p.ng2 = (((K + CL - 1) / CL) + 63) / 64 * 64;
  p.store_t = t_bf16 != nullptr;
  p.ap = ap; p.ldap = ldap; p.apb = reinterpret_cast<__nv_bfloat16*>(ap_bf16); p.ldapb = ldapb;
  p.nclusters = 0;
  return CL == 2 ? launch<2>(maps, p, st) : launch<1>(maps, p, st);
}

}  // namespace chain
}  // namespace gp

extern "C" int gp_pool_chain_bf16(const void* s_bf16, long long lds, const void* adj_bf16, long long ldadj,
                                  const int32_t* nb, const int32_t* order, int B, int N, int K, void* t_bf16,
                                  long long ldt, float* ap, long long ldap, void* ap_bf16, long long ldapb,
                                  gp_stream_t stream) {
  return gp::chain::run(s_bf16, lds, adj_bf16, ldadj, nb, order, B, N, K, t_bf16, ldt, ap, ldap, ap_bf16, ldapb,
                        gp::S(stream));
}
