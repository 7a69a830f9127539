// Vectorised single-pass backward of one GCN layer's element-wise tail (HBM-bound):
//   [concat-slot gradient + next layer's dX + max-readout scatter] -> BatchNorm-per-node-index (batch
//   statistics) -> ReLU -> L2-normalize   (encoders.py:1062-1064,1048-1052,323-326)
// producing dV in fp32 and/or bf16 (the operand of the dW / dU contractions) and the column sums of dV
// (= the bias gradient) in the same pass, so every input is read from HBM exactly once.
//
// Two kernels:
//  * layer_bwd_bn_kernel   BN couples all graphs at one node index n.  A thread-block CLUSTER of CS CTAs owns
//    node n; each CTA keeps its B/CS rows in registers (16-byte loads, everything in flight at once), the
//    cluster reduces mean(g) and mean(g*Hhat) through distributed shared memory, then each CTA finishes its
//    rows.  Several clusters are resident per SM, so one node's reduction overlaps another node's loads.
//  * layer_bwd_row_kernel  no BN (last layer of a stack): rows are independent, one warp per row.
// Shapes outside the fast paths (d % 4 != 0, unaligned strides, d > 512, ...) use rowops.cu's generic kernel.
#include <cooperative_groups.h>
#include <stdlib.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace gp {

constexpr float kEpsNormB = 1e-12f;

int colsum(const float* x, long long rows, int d, long long ld, float* out, int accumulate, float* ws,
           cudaStream_t st);
int layer_bwd_generic(const gp_layer_bwd* a, cudaStream_t st);   // rowops.cu

struct LbArgs {
  const float* dz; long long lddz;
  const float* dxn; long long lddxn;
  const float* dout; const int32_t* argidx; long long ldo;
  const float* h; long long ldh;
  const float* y; long long ldy;
  const float* rnorm; const float* mean; const float* invstd;
  int B, N, d, relu, bn, normalize;
  float* dv; __nv_bfloat16* dvb; long long lddvb;
  float* part;          // [blocks][d] partial column sums or NULL
  const int32_t* nbz;   // rows n >= nbz[b] have zero upstream gradient (row kernels: dV = 0 written without reading)
  int dz_bf16, dxn_bf16;   // dz / dxn point to bf16 data (strides in elements): gradient intermediates of the bf16 schedule
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// read-once operands: streaming (evict-first) loads keep them from displacing the Y rows the kernel reads twice
__device__ __forceinline__ float4 ld4s(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }

// 4 consecutive bf16 values as floats (8-byte load)
template <bool STREAM>
__device__ __forceinline__ float4 ld4h(const __nv_bfloat16* p) {
  const uint2 w = STREAM ? __ldcs(reinterpret_cast<const uint2*>(p)) : *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u), __uint_as_float(w.y << 16),
                     __uint_as_float(w.y & 0xffff0000u));
}

// combined upstream gradient of 4 consecutive columns of row (b, n)
template <bool STREAM = false>
__device__ __forceinline__ float4 load_g(const LbArgs& a, int b, int n, long long row, int c) {
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.dz != nullptr) {
    if (a.dz_bf16) g = ld4h<STREAM>(reinterpret_cast<const __nv_bfloat16*>(a.dz) + row * a.lddz + c);
    else g = STREAM ? ld4s(a.dz + row * a.lddz + c) : ld4(a.dz + row * a.lddz + c);
  }
  if (a.dxn != nullptr) {
    float4 t;
    if (a.dxn_bf16) t = ld4h<STREAM>(reinterpret_cast<const __nv_bfloat16*>(a.dxn) + row * a.lddxn + c);
    else t = STREAM ? ld4s(a.dxn + row * a.lddxn + c) : ld4(a.dxn + row * a.lddxn + c);
    g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
  }
  if (a.dout != nullptr) {
    const int4 i4 = *reinterpret_cast<const int4*>(a.argidx + (long long)b * a.ldo + c);
    const float4 o = ld4(a.dout + (long long)b * a.ldo + c);
    if (i4.x == n) g.x += o.x;
    if (i4.y == n) g.y += o.y;
    if (i4.z == n) g.z += o.z;
    if (i4.w == n) g.w += o.w;
  }
  return g;
}

// The same with the source configuration known at compile time (CFG >= 0: dz and dxn both present; bit 0: dz is bf16,
// bit 1: dxn is bf16, bit 2: the readout scatter is present): no per-pass branches, so the loads of several passes can
// be issued back to back.  CFG < 0: the run-time form above.
template <bool STREAM, int CFG>
__device__ __forceinline__ float4 load_g_cfg(const LbArgs& a, int b, int n, long long row, int c) {
  if constexpr (CFG < 0) {
    return load_g<STREAM>(a, b, n, row, c);
  } else {
    float4 g, t;
    if constexpr ((CFG & 1) != 0) g = ld4h<STREAM>(reinterpret_cast<const __nv_bfloat16*>(a.dz) + row * a.lddz + c);
    else g = STREAM ? ld4s(a.dz + row * a.lddz + c) : ld4(a.dz + row * a.lddz + c);
    if constexpr ((CFG & 2) != 0) t = ld4h<STREAM>(reinterpret_cast<const __nv_bfloat16*>(a.dxn) + row * a.lddxn + c);
    else t = STREAM ? ld4s(a.dxn + row * a.lddxn + c) : ld4(a.dxn + row * a.lddxn + c);
    g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
    if constexpr ((CFG & 4) != 0) {
      const int4 i4 = *reinterpret_cast<const int4*>(a.argidx + (long long)b * a.ldo + c);
      const float4 o = ld4(a.dout + (long long)b * a.ldo + c);
      if (i4.x == n) g.x += o.x;
      if (i4.y == n) g.y += o.y;
      if (i4.z == n) g.z += o.z;
      if (i4.w == n) g.w += o.w;
    }
    return g;
  }
}

__device__ __forceinline__ void store_dv(const LbArgs& a, long long row, int c, float4 v) {
  if (a.dv != nullptr) *reinterpret_cast<float4*>(a.dv + row * a.d + c) = v;
  if (a.dvb != nullptr)
    *reinterpret_cast<uint2*>(a.dvb + row * a.lddvb + c) = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
}

// ---------------------------------------------------------------------------------------------------------
// BN layers.  T threads, D4 = d/4 (power of two <= 32) threads per row, VPT rows per thread.
// ---------------------------------------------------------------------------------------------------------
template <int VPT, bool HAS_H>
__global__ void __launch_bounds__(256) layer_bwd_bn_kernel(const LbArgs a, int CS, int rows_per_cta) {
  constexpr int T = 256;
  __shared__ float red[2][T / 32];
  __shared__ float cta_part[2];
  __shared__ __align__(16) float colacc[T * 4];
  cg::cluster_group cluster = cg::this_cluster();
  const int n = blockIdx.x / CS;
  const int rank = blockIdx.x - n * CS;
  const int d4 = a.d >> 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = (tid % d4) * 4;
  const int rstep = T / d4;
  const int b0 = rank * rows_per_cta;
  const int b1 = min(a.B, b0 + rows_per_cta);
  const float mu = a.mean != nullptr ? a.mean[n] : 0.f;
  const float is = a.invstd[n];

  float4 g[VPT], yv[VPT], hh[HAS_H ? VPT : 1];
  float rn[VPT];
  // ---- all loads in flight ----
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const int b = b0 + tid / d4 + j * rstep;
    g[j] = yv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (HAS_H) hh[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    rn[j] = 1.f;
    if (b < b1) {
      const long long row = (long long)b * a.N + n;
      g[j] = load_g(a, b, n, row, c);
      yv[j] = ld4(a.y + row * a.ldy + c);
      if (HAS_H) hh[j] = ld4(a.h + row * a.ldh + c);
      if (a.normalize) rn[j] = a.rnorm[row];
    }
  }
  auto hhat = [&](int j) -> float4 {                     // stored BN output, or recomputed from Y and the statistics
    if (HAS_H) return hh[j];
    const float4 y = yv[j];
    return make_float4(((a.relu ? fmaxf(y.x, 0.f) : y.x) - mu) * is, ((a.relu ? fmaxf(y.y, 0.f) : y.y) - mu) * is,
                       ((a.relu ? fmaxf(y.z, 0.f) : y.z) - mu) * is, ((a.relu ? fmaxf(y.w, 0.f) : y.w) - mu) * is);
  };
  // ---- Hhat and the two batch means ----
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const float4 hj = hhat(j);                           // rows beyond b1 have g == 0: they add nothing
    s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
    s2 = fmaf(g[j].x, hj.x, s2); s2 = fmaf(g[j].y, hj.y, s2);
    s2 = fmaf(g[j].z, hj.z, s2); s2 = fmaf(g[j].w, hj.w, s2);
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
  __syncthreads();
  if (tid == 0) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int w = 0; w < T / 32; ++w) { t1 += red[0][w]; t2 += red[1][w]; }
    cta_part[0] = t1; cta_part[1] = t2;
  }
  cluster.sync();                                        // every CTA's partial is visible cluster-wide
  float S1 = 0.f, S2 = 0.f;
  for (int r = 0; r < CS; ++r) {                         // same order in every CTA: bitwise-identical means
    const float* rp = cluster.map_shared_rank(cta_part, r);
    S1 += rp[0]; S2 += rp[1];
  }
  cluster.sync();                                        // nobody exits while its partial may still be read
  const float inv_cnt = 1.f / ((float)a.B * (float)a.d);
  const float m1 = S1 * inv_cnt, m2 = S2 * inv_cnt;

  // ---- finish the rows ----
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const int b = b0 + tid / d4 + j * rstep;
    const float4 hj = hhat(j);
    float4 v;
    v.x = (g[j].x - m1 - hj.x * m2) * is;
    v.y = (g[j].y - m1 - hj.y * m2) * is;
    v.z = (g[j].z - m1 - hj.z * m2) * is;
    v.w = (g[j].w - m1 - hj.w * m2) * is;
    const float4 y = yv[j];
    if (a.relu) {
      if (!(y.x > 0.f)) v.x = 0.f;
      if (!(y.y > 0.f)) v.y = 0.f;
      if (!(y.z > 0.f)) v.z = 0.f;
      if (!(y.w > 0.f)) v.w = 0.f;
    }
    if (a.normalize) {
      float dot = fmaf(v.x, y.x, fmaf(v.y, y.y, fmaf(v.z, y.z, v.w * y.w)));
      for (int o = d4 >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      const float r = rn[j];
      if (!(r > kEpsNormB)) {
        v.x /= kEpsNormB; v.y /= kEpsNormB; v.z /= kEpsNormB; v.w /= kEpsNormB;
      } else {
        const float ir = 1.f / r;
        v.x = (v.x - y.x * dot) * ir; v.y = (v.y - y.y * dot) * ir;
        v.z = (v.z - y.z * dot) * ir; v.w = (v.w - y.w * dot) * ir;
      }
    }
    if (b < b1) {
      store_dv(a, (long long)b * a.N + n, c, v);
      cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
    }
  }
  if (a.part != nullptr) {                               // deterministic per-CTA column sums
    *reinterpret_cast<float4*>(&colacc[tid * 4]) = cs;
    __syncthreads();
    if (tid < a.d) {
      const int q = tid >> 2, e = tid & 3;               // column tid = 4*q + e lives in threads q, q+d4, ...
      float t = 0.f;
      for (int r = q; r < T; r += d4) t += colacc[r * 4 + e];
      a.part[(long long)blockIdx.x * a.d + tid] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// One-CTA-per-node BN variant (no cluster).  Cluster barriers and distributed-shared-memory exchanges cost
// microseconds (measured: they, not HBM, bound the cluster kernel at 2.5 TB/s).  When a node's whole batch fits
// ONE 1024-thread CTA (B * d <= 32768 elements at 8 float4 per thread), the two batch means are a plain block
// reduction: phase 1 loads every gradient source and Y with everything in flight (up to 384 KB per CTA), keeps only
// g in registers; phase 2 re-reads Y (L2-resident: the CTA just read it) and finishes the rows.
// ---------------------------------------------------------------------------------------------------------
template <int VPT, bool STREAM = false>
__global__ void __launch_bounds__(1024, 1) layer_bwd_bn_cta_kernel(const LbArgs a) {
  constexpr int T = 1024;
  __shared__ float red[2][T / 32];
  __shared__ float tot[2];
  __shared__ __align__(16) float colacc[T * 4];
  const int n = blockIdx.x;
  const int d4 = a.d >> 2;
  const int lg4 = 31 - __clz(d4);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = (tid & (d4 - 1)) * 4;
  const int rstep = T >> lg4;
  const int r0 = tid >> lg4;
  const float mu = a.mean[n], is = a.invstd[n];
  auto hhat4 = [&](const float4 y) -> float4 {
    return make_float4(((a.relu ? fmaxf(y.x, 0.f) : y.x) - mu) * is, ((a.relu ? fmaxf(y.y, 0.f) : y.y) - mu) * is,
                       ((a.relu ? fmaxf(y.z, 0.f) : y.z) - mu) * is, ((a.relu ? fmaxf(y.w, 0.f) : y.w) - mu) * is);
  };
  float4 g[VPT];
  float s1 = 0.f, s2 = 0.f;
  {
    float4 yv[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) {                      // ---- all loads in flight ----
      const int b = r0 + j * rstep;
      g[j] = yv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b < a.B) {
        const long long row = (long long)b * a.N + n;
        g[j] = load_g<STREAM>(a, b, n, row, c);
        yv[j] = ld4(a.y + row * a.ldy + c);
      }
    }
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const float4 hj = hhat4(yv[j]);                    // rows beyond B have g == 0: they add nothing
      s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
      s2 = fmaf(g[j].x, hj.x, s2); s2 = fmaf(g[j].y, hj.y, s2);
      s2 = fmaf(g[j].z, hj.z, s2); s2 = fmaf(g[j].w, hj.w, s2);
    }
  }
  // phase 2 operands: issue the Y / rnorm re-reads before the reduction so their latency overlaps it
  float4 y2[VPT];
  float rn[VPT];
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const int b = r0 + j * rstep;
    y2[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    rn[j] = 1.f;
    if (b < a.B) {
      const long long row = (long long)b * a.N + n;
      y2[j] = ld4(a.y + row * a.ldy + c);
      if (a.normalize) rn[j] = a.rnorm[row];
    }
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
  __syncthreads();
  if (warp == 0) {
    float t1 = red[0][lane], t2 = red[1][lane];          // T / 32 == 32 partials
    t1 = warp_sum(t1); t2 = warp_sum(t2);
    if (lane == 0) { tot[0] = t1; tot[1] = t2; }
  }
  __syncthreads();
  const float inv_cnt = 1.f / ((float)a.B * (float)a.d);
  const float m1 = tot[0] * inv_cnt, m2 = tot[1] * inv_cnt;
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const int b = r0 + j * rstep;
    const float4 y = y2[j];
    const float4 hj = hhat4(y);
    float4 v;
    v.x = (g[j].x - m1 - hj.x * m2) * is; v.y = (g[j].y - m1 - hj.y * m2) * is;
    v.z = (g[j].z - m1 - hj.z * m2) * is; v.w = (g[j].w - m1 - hj.w * m2) * is;
    if (a.relu) {
      if (!(y.x > 0.f)) v.x = 0.f;
      if (!(y.y > 0.f)) v.y = 0.f;
      if (!(y.z > 0.f)) v.z = 0.f;
      if (!(y.w > 0.f)) v.w = 0.f;
    }
    if (a.normalize) {
      float dot = fmaf(v.x, y.x, fmaf(v.y, y.y, fmaf(v.z, y.z, v.w * y.w)));
      for (int o = d4 >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      const float r = rn[j];
      if (!(r > kEpsNormB)) {
        v.x /= kEpsNormB; v.y /= kEpsNormB; v.z /= kEpsNormB; v.w /= kEpsNormB;
      } else {
        const float ir = 1.f / r;
        v.x = (v.x - y.x * dot) * ir; v.y = (v.y - y.y * dot) * ir;
        v.z = (v.z - y.z * dot) * ir; v.w = (v.w - y.w * dot) * ir;
      }
    }
    if (b < a.B) {
      store_dv(a, (long long)b * a.N + n, c, v);
      cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
    }
  }
  if (a.part != nullptr) {                               // deterministic per-CTA column sums
    *reinterpret_cast<float4*>(&colacc[tid * 4]) = cs;
    __syncthreads();
    if (tid < a.d) {
      const int q = tid >> 2, e = tid & 3;
      float t = 0.f;
      for (int r = q; r < T; r += d4) t += colacc[r * 4 + e];
      a.part[(long long)blockIdx.x * a.d + tid] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Two-CTAs-per-SM BN variant (bf16-only output, i.e. the tensor-core schedule).  ncu on the kernel above
// (profiles/r1_ncu_layer_bwd_bn_cta_n2048.md): DRAM 39 % busy, one CTA resident per SM, so the load phase of one node
// never overlaps the reduce / finish / store phase of another.  Here a node is one 512-thread CTA with twice the rows
// per thread; the combined upstream gradient g is stashed between the two phases as PACKED BF16 (2 registers per
// float4 instead of 4), which brings a thread under the 64-register line needed for two resident CTAs.  The batch sums
// s1 = sum g, s2 = sum g.hhat are taken from the fp32 values BEFORE the stash; only the per-element use of g in phase 2
// sees the rounding, of the same size as the bf16 rounding of the dV it produces (this variant is used only when no
// fp32 dV is requested).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// FULL: B == VPT * rows-per-pass, so no row of any pass falls outside the batch -- straight-line passes without bounds
// branches, which also lets the compiler batch the loads of several passes (the kernel waits on load latency: ncu r2)
// CFG >= 0: gradient sources fixed at compile time (load_g_cfg) and ReLU + normalize both on (every BatchNorm layer of
// the tensor-core schedule): branch-free passes.
template <int VPT, bool FULL = false, int CFG = -1>
__global__ void __launch_bounds__(512, 2) layer_bwd_bn_cta2_kernel(const LbArgs a) {
  constexpr bool FIX = CFG >= 0;
  const bool relu = FIX ? true : (a.relu != 0), normalize = FIX ? true : (a.normalize != 0);
  constexpr int T = 512;
  __shared__ float red[2][T / 32];
  __shared__ float tot[2];
  __shared__ __align__(16) float colacc[T * 4];
  const int n = blockIdx.x;
  const int d4 = a.d >> 2;
  const int lg4 = 31 - __clz(d4);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = (tid & (d4 - 1)) * 4;
  const int rstep = T >> lg4;
  const int r0 = tid >> lg4;
  const float mu = a.mean[n], is = a.invstd[n];
  auto hhat4 = [&](const float4 y) -> float4 {
    return make_float4(((relu ? fmaxf(y.x, 0.f) : y.x) - mu) * is, ((relu ? fmaxf(y.y, 0.f) : y.y) - mu) * is,
                       ((relu ? fmaxf(y.z, 0.f) : y.z) - mu) * is, ((relu ? fmaxf(y.w, 0.f) : y.w) - mu) * is);
  };
  uint2 gs[VPT];                                         // g as 4 x bf16
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const int b = r0 + j * rstep;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f), yv = g;
    if (FULL || b < a.B) {
      const long long row = (long long)b * a.N + n;
      g = load_g_cfg<true, CFG>(a, b, n, row, c);
      yv = ld4(a.y + row * a.ldy + c);
    }
    const float4 hj = hhat4(yv);                         // rows beyond B have g == 0: they add nothing
    s1 += (g.x + g.y) + (g.z + g.w);
    s2 = fmaf(g.x, hj.x, s2); s2 = fmaf(g.y, hj.y, s2);
    s2 = fmaf(g.z, hj.z, s2); s2 = fmaf(g.w, hj.w, s2);
    gs[j] = make_uint2(pack2(g.x, g.y), pack2(g.z, g.w));
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
  __syncthreads();
  if (warp == 0) {
    float t1 = lane < T / 32 ? red[0][lane] : 0.f, t2 = lane < T / 32 ? red[1][lane] : 0.f;
    t1 = warp_sum(t1); t2 = warp_sum(t2);
    if (lane == 0) { tot[0] = t1; tot[1] = t2; }
  }
  __syncthreads();
  const float inv_cnt = 1.f / ((float)a.B * (float)a.d);
  const float m1 = tot[0] * inv_cnt, m2 = tot[1] * inv_cnt;
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const int b = r0 + j * rstep;
    const bool live = FULL || b < a.B;
    const long long row = (long long)(live ? b : 0) * a.N + n;
    float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
    float r = 1.f;
    if (live) {
      y = ld4(a.y + row * a.ldy + c);                    // second read of Y: L2 (this CTA just read it)
      if (normalize) r = a.rnorm[row];
    }
    const float4 hj = hhat4(y);
    const float4 g = make_float4(bf_lo(gs[j].x), bf_hi(gs[j].x), bf_lo(gs[j].y), bf_hi(gs[j].y));
    float4 v;
    v.x = (g.x - m1 - hj.x * m2) * is; v.y = (g.y - m1 - hj.y * m2) * is;
    v.z = (g.z - m1 - hj.z * m2) * is; v.w = (g.w - m1 - hj.w * m2) * is;
    if (relu) {
      if (!(y.x > 0.f)) v.x = 0.f;
      if (!(y.y > 0.f)) v.y = 0.f;
      if (!(y.z > 0.f)) v.z = 0.f;
      if (!(y.w > 0.f)) v.w = 0.f;
    }
    if (normalize) {
      float dot = fmaf(v.x, y.x, fmaf(v.y, y.y, fmaf(v.z, y.z, v.w * y.w)));
      for (int o = d4 >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      if (!(r > kEpsNormB)) {
        v.x /= kEpsNormB; v.y /= kEpsNormB; v.z /= kEpsNormB; v.w /= kEpsNormB;
      } else {
        const float ir = 1.f / r;
        v.x = (v.x - y.x * dot) * ir; v.y = (v.y - y.y * dot) * ir;
        v.z = (v.z - y.z * dot) * ir; v.w = (v.w - y.w * dot) * ir;
      }
    }
    if (live) {
      *reinterpret_cast<uint2*>(a.dvb + row * a.lddvb + c) = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
      cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
    }
  }
  if (a.part != nullptr) {                               // deterministic per-CTA column sums
    *reinterpret_cast<float4*>(&colacc[tid * 4]) = cs;
    __syncthreads();
    if (tid < a.d) {
      const int q = tid >> 2, e = tid & 3;
      float t = 0.f;
      for (int rr = q; rr < T; rr += d4) t += colacc[rr * 4 + e];
      a.part[(long long)blockIdx.x * a.d + tid] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Pipelined BN variant: PERSISTENT clusters.  The one-node-per-cluster kernel above spends most of a CTA's life
// outside its load phase (cluster syncs, reduction, launch of the next cluster), so HBM idles: measured 2.5 TB/s.
// Here a cluster of CS CTAs walks nodes n = cluster, cluster + G, ...; the rows of node i+1 stream into the other
// shared-memory stage with cp.async (16 B per thread) while node i is reduced and finished, so loads are always in
// flight.  One cluster.sync per node (the exchanged partial is double-buffered by iteration parity).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NKEEP>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NKEEP) : "memory"); }

template <int T, int NST>
__global__ void __launch_bounds__(T, 1)
layer_bwd_bn_pipe_kernel(const LbArgs a, int CS, int rows_per_cta, int nclusters) {
  extern __shared__ __align__(16) float smem_f[];
  __shared__ float red[2][T / 32];
  __shared__ float cta_part[2][2];
  __shared__ __align__(16) float colacc[T * 4];
  cg::cluster_group cluster = cg::this_cluster();
  const int cl = blockIdx.x / CS;
  const int rank = blockIdx.x - cl * CS;
  const int d = a.d, d4 = d >> 2;
  const int lg4 = 31 - __clz(d4);                        // d4 is a power of two: divisions become shifts
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = (tid & (d4 - 1)) * 4;
  const int rstep = T >> lg4;
  const int b0 = rank * rows_per_cta;
  const int b1 = min(a.B, b0 + rows_per_cta);
  const int nrows = max(b1 - b0, 0);
  const bool has_dz = a.dz != nullptr, has_dx = a.dxn != nullptr;
  const int nstreams = 1 + (has_dz ? 1 : 0) + (has_dx ? 1 : 0);
  const int plane = rows_per_cta * d;                    // floats per stream per stage
  const int rn_f = (rows_per_cta + 3) & ~3;              // per-row divisors of the L2 normalisation ride along
  const int stage_f = plane * nstreams + rn_f;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_f);

  auto issue = [&](int n, int stage) {                   // rows of node n -> stage (planes: y, dz, dxn)
    const uint32_t sb = smem_base + (uint32_t)(stage * stage_f) * 4u;
    const int per = nrows * d4;
    for (int idx = tid; idx < per; idx += T) {
      const int r = idx >> lg4, c4 = idx & (d4 - 1);
      const long long row = (long long)(b0 + r) * a.N + n;
      const uint32_t off = (uint32_t)(r * d + c4 * 4) * 4u;
      cp_async16(sb + off, a.y + row * a.ldy + c4 * 4);
      int pl = 1;
      if (has_dz) { cp_async16(sb + (uint32_t)(pl * plane) * 4u + off, a.dz + row * a.lddz + c4 * 4); ++pl; }
      if (has_dx) { cp_async16(sb + (uint32_t)(pl * plane) * 4u + off, a.dxn + row * a.lddxn + c4 * 4); }
    }
    if (a.normalize)
      for (int r = tid; r < nrows; r += T)
        cp_async4(sb + (uint32_t)(plane * nstreams + r) * 4u, a.rnorm + (long long)(b0 + r) * a.N + n);
    cp_async_commit();
  };

  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  const float inv_cnt = 1.f / ((float)a.B * (float)d);
  int it = 0;
  int n = cl;
  // prologue: NST-1 nodes in flight (empty groups keep the wait_group arithmetic uniform at the tail)
#pragma unroll
  for (int j = 0; j < NST - 1; ++j) {
    const int nj = cl + j * nclusters;
    if (nj < a.N) issue(nj, j); else cp_async_commit();
  }
  for (; n < a.N; n += nclusters, ++it) {
    const int stage = it % NST;
    const int nn = n + (NST - 1) * nclusters;
    const float mu = a.mean[n], is = a.invstd[n];        // issued before the wait below: latency overlaps it
    if (nn < a.N) issue(nn, (it + NST - 1) % NST); else cp_async_commit();
    cp_async_wait<NST - 1>();
    __syncthreads();
    float* sy = smem_f + stage * stage_f;
    float* sg = sy + plane;                              // plane 1 holds dz (or dxn if no dz); becomes g in place
    float* sx = sy + 2 * plane;
    const float* srn = sy + plane * nstreams;
    // ---- pass 1: g (kept in shared memory), the two batch sums ----
    float s1 = 0.f, s2 = 0.f;
    for (int r = tid >> lg4; r < nrows; r += rstep) {     // (no warp collectives inside: lanes may drop out)
      const int b = b0 + r;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (nstreams >= 2) g = *reinterpret_cast<const float4*>(sg + r * d + c);
      if (nstreams == 3) {
        const float4 t = *reinterpret_cast<const float4*>(sx + r * d + c);
        g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
      }
      if (a.dout != nullptr) {
        const int4 i4 = *reinterpret_cast<const int4*>(a.argidx + (long long)b * a.ldo + c);
        const float4 o = ld4(a.dout + (long long)b * a.ldo + c);
        if (i4.x == n) g.x += o.x;
        if (i4.y == n) g.y += o.y;
        if (i4.z == n) g.z += o.z;
        if (i4.w == n) g.w += o.w;
      }
      const float4 y = *reinterpret_cast<const float4*>(sy + r * d + c);
      const float hx = ((a.relu ? fmaxf(y.x, 0.f) : y.x) - mu) * is, hy = ((a.relu ? fmaxf(y.y, 0.f) : y.y) - mu) * is;
      const float hz = ((a.relu ? fmaxf(y.z, 0.f) : y.z) - mu) * is, hw = ((a.relu ? fmaxf(y.w, 0.f) : y.w) - mu) * is;
      s1 += (g.x + g.y) + (g.z + g.w);
      s2 = fmaf(g.x, hx, s2); s2 = fmaf(g.y, hy, s2); s2 = fmaf(g.z, hz, s2); s2 = fmaf(g.w, hw, s2);
      if (nstreams >= 2) *reinterpret_cast<float4*>(sg + r * d + c) = g;   // same thread re-reads it in pass 2
      else if (a.dout != nullptr) { /* g lives only in the scatter: recomputed in pass 2 */ }
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
    __syncthreads();
    if (tid == 0) {
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int w = 0; w < T / 32; ++w) { t1 += red[0][w]; t2 += red[1][w]; }
      cta_part[it & 1][0] = t1; cta_part[it & 1][1] = t2;
    }
    cluster.sync();
    float S1 = 0.f, S2 = 0.f;
    for (int r = 0; r < CS; ++r) {                       // same order in every CTA: bitwise-identical means
      const float* rp = cluster.map_shared_rank(&cta_part[it & 1][0], r);
      S1 += rp[0]; S2 += rp[1];
    }
    const float m1 = S1 * inv_cnt, m2 = S2 * inv_cnt;
    // ---- pass 2: finish the rows (block-uniform trip count: the row's dot product is a warp shuffle) ----
    for (int r0 = 0; r0 < nrows; r0 += rstep) {
      const int r = r0 + (tid >> lg4);
      const bool live = r < nrows;
      const int b = b0 + r;
      const long long row = (long long)b * a.N + n;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!live) {
        // nothing to load: g = y = 0
      } else if (nstreams >= 2) {
        g = *reinterpret_cast<const float4*>(sg + r * d + c);
      } else if (a.dout != nullptr) {
        const int4 i4 = *reinterpret_cast<const int4*>(a.argidx + (long long)b * a.ldo + c);
        const float4 o = ld4(a.dout + (long long)b * a.ldo + c);
        if (i4.x == n) g.x += o.x;
        if (i4.y == n) g.y += o.y;
        if (i4.z == n) g.z += o.z;
        if (i4.w == n) g.w += o.w;
      }
      const float4 y = live ? *reinterpret_cast<const float4*>(sy + r * d + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float hx = ((a.relu ? fmaxf(y.x, 0.f) : y.x) - mu) * is, hy = ((a.relu ? fmaxf(y.y, 0.f) : y.y) - mu) * is;
      const float hz = ((a.relu ? fmaxf(y.z, 0.f) : y.z) - mu) * is, hw = ((a.relu ? fmaxf(y.w, 0.f) : y.w) - mu) * is;
      float4 v;
      v.x = (g.x - m1 - hx * m2) * is; v.y = (g.y - m1 - hy * m2) * is;
      v.z = (g.z - m1 - hz * m2) * is; v.w = (g.w - m1 - hw * m2) * is;
      if (a.relu) {
        if (!(y.x > 0.f)) v.x = 0.f;
        if (!(y.y > 0.f)) v.y = 0.f;
        if (!(y.z > 0.f)) v.z = 0.f;
        if (!(y.w > 0.f)) v.w = 0.f;
      }
      if (a.normalize) {
        float dot = fmaf(v.x, y.x, fmaf(v.y, y.y, fmaf(v.z, y.z, v.w * y.w)));
        for (int o = d4 >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        const float rr = live ? srn[r] : 1.f;
        if (!(rr > kEpsNormB)) {
          v.x /= kEpsNormB; v.y /= kEpsNormB; v.z /= kEpsNormB; v.w /= kEpsNormB;
        } else {
          const float ir = 1.f / rr;
          v.x = (v.x - y.x * dot) * ir; v.y = (v.y - y.y * dot) * ir;
          v.z = (v.z - y.z * dot) * ir; v.w = (v.w - y.w * dot) * ir;
        }
      }
      if (live) {
        store_dv(a, row, c, v);
        cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
      }
    }
    __syncthreads();                                     // this stage may be refilled two iterations from now
  }
  if (a.part != nullptr) {                               // deterministic per-CTA column sums over all its nodes
    *reinterpret_cast<float4*>(&colacc[tid * 4]) = cs;
    __syncthreads();
    if (tid < d) {
      const int q = tid >> 2, e = tid & 3;
      float t = 0.f;
      for (int r = q; r < T; r += d4) t += colacc[r * 4 + e];
      a.part[(long long)blockIdx.x * d + tid] = t;
    }
  }
  cluster.sync();                                        // nobody exits while its partials may still be read
}

// ---------------------------------------------------------------------------------------------------------
// No BN: one warp per row, lane owns float4 columns lane, lane+32, ... (VPL of them).
// ---------------------------------------------------------------------------------------------------------
// CFG >= 0 (the last layer of a stack in the tensor-core schedule): dz present (bit 0: bf16), no dxn, bit 2: readout
// scatter present, no ReLU, normalize on, row width exactly 128 * VPL floats -- no per-element branches.
template <int VPL, int MINB = 1, int CFG = -1>
__global__ void __launch_bounds__(256, MINB) layer_bwd_row_kernel(const LbArgs a, long long rows) {
  constexpr bool FIX = CFG >= 0;
  const bool relu = FIX ? false : (a.relu != 0), normalize = FIX ? true : (a.normalize != 0);
  __shared__ __align__(16) float colacc[8][VPL * 128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d4 = a.d >> 2;
  const long long gw = (long long)blockIdx.x * 8 + warp, nw = (long long)gridDim.x * 8;
  float4 cs[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) cs[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long row = gw; row < rows; row += nw) {
    int b, n;
    row_split(row, a.N, b, n);
    if (a.nbz != nullptr && n >= a.nbz[b]) {             // pad row of a masked level: dV = 0, nothing to read
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int c4 = lane + 32 * k;
        if (c4 < d4) store_dv(a, row, c4 * 4, make_float4(0.f, 0.f, 0.f, 0.f));
      }
      continue;
    }
    float4 g[VPL], yv[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c4 = lane + 32 * k;
      g[k] = yv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (FIX || c4 < d4) {
        if constexpr (FIX) {                             // every operand of a row is read exactly once: streaming
          if constexpr ((CFG & 1) != 0) g[k] = ld4h<true>(reinterpret_cast<const __nv_bfloat16*>(a.dz) + row * a.lddz + c4 * 4);
          else g[k] = ld4s(a.dz + row * a.lddz + c4 * 4);
          if constexpr ((CFG & 4) != 0) {
            const int4 i4 = *reinterpret_cast<const int4*>(a.argidx + (long long)b * a.ldo + c4 * 4);
            const float4 o = ld4(a.dout + (long long)b * a.ldo + c4 * 4);
            if (i4.x == n) g[k].x += o.x;
            if (i4.y == n) g[k].y += o.y;
            if (i4.z == n) g[k].z += o.z;
            if (i4.w == n) g[k].w += o.w;
          }
        } else {
          g[k] = load_g<true>(a, b, n, row, c4 * 4);
        }
        yv[k] = ld4s(a.y + row * a.ldy + c4 * 4);
      }
    }
    const float r = normalize ? a.rnorm[row] : 1.f;
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (relu) {
        if (!(yv[k].x > 0.f)) g[k].x = 0.f;
        if (!(yv[k].y > 0.f)) g[k].y = 0.f;
        if (!(yv[k].z > 0.f)) g[k].z = 0.f;
        if (!(yv[k].w > 0.f)) g[k].w = 0.f;
      }
      dot = fmaf(g[k].x, yv[k].x, fmaf(g[k].y, yv[k].y, fmaf(g[k].z, yv[k].z, fmaf(g[k].w, yv[k].w, dot))));
    }
    if (normalize) {
      dot = warp_sum(dot);
      const bool clamped = !(r > kEpsNormB);
      const float ir = clamped ? 1.f / kEpsNormB : 1.f / r;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        if (clamped) {
          g[k].x /= kEpsNormB; g[k].y /= kEpsNormB; g[k].z /= kEpsNormB; g[k].w /= kEpsNormB;
        } else {
          g[k].x = (g[k].x - yv[k].x * dot) * ir; g[k].y = (g[k].y - yv[k].y * dot) * ir;
          g[k].z = (g[k].z - yv[k].z * dot) * ir; g[k].w = (g[k].w - yv[k].w * dot) * ir;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c4 = lane + 32 * k;
      if (FIX || c4 < d4) {
        store_dv(a, row, c4 * 4, g[k]);
        cs[k].x += g[k].x; cs[k].y += g[k].y; cs[k].z += g[k].z; cs[k].w += g[k].w;
      }
    }
  }
  if (a.part != nullptr) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) *reinterpret_cast<float4*>(&colacc[warp][(lane + 32 * k) * 4]) = cs[k];
    __syncthreads();
    for (int cidx = threadIdx.x; cidx < a.d; cidx += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += colacc[w][cidx];
      a.part[(long long)blockIdx.x * a.d + cidx] = t;
    }
  }
}

// Rows wider than 512 floats (the assignment GCN's last layer at K = 1250 -> 1256 / 1280 clusters): the row no longer
// fits the registers twice over, so the dot product is taken in a first pass and the operands are read again (L1 / L2
// hits: a row is a few KB) for the update.  4 warps per block; VPL float4 per lane hold the column sums only.
// CFG >= 0: dz is the only gradient source (CFG 0: fp32, 1: bf16), no ReLU, normalize on (the assignment GCN's last layer)
template <int VPL, int MINB = 1, int CFG = -1>
__global__ void __launch_bounds__(128, MINB) layer_bwd_row_wide_kernel(const LbArgs a, long long rows) {
  constexpr bool FIX = CFG >= 0;
  const bool relu = FIX ? false : (a.relu != 0), normalize = FIX ? true : (a.normalize != 0);
  auto ldg = [&](int b, int n, long long row, int c) -> float4 {
    if constexpr (CFG == 0) return ld4(a.dz + row * a.lddz + c);
    else if constexpr (CFG == 1) return ld4h<false>(reinterpret_cast<const __nv_bfloat16*>(a.dz) + row * a.lddz + c);
    else return load_g(a, b, n, row, c);
  };
  __shared__ __align__(16) float colacc[4][VPL * 128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d4 = a.d >> 2;
  const long long gw = (long long)blockIdx.x * 4 + warp, nw = (long long)gridDim.x * 4;
  float4 cs[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) cs[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long row = gw; row < rows; row += nw) {
    int b, n;
    row_split(row, a.N, b, n);
    if (a.nbz != nullptr && n >= a.nbz[b]) {             // pad row of a masked level: dV = 0, nothing to read
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int c4 = lane + 32 * k;
        if (c4 < d4) store_dv(a, row, c4 * 4, make_float4(0.f, 0.f, 0.f, 0.f));
      }
      continue;
    }
    const float r = normalize ? a.rnorm[row] : 1.f;
    float dot = 0.f;
    if (normalize) {
#pragma unroll 4
      for (int k = 0; k < VPL; ++k) {
        const int c4 = lane + 32 * k;
        if (c4 < d4) {
          float4 g = ldg(b, n, row, c4 * 4);
          const float4 yv = ld4(a.y + row * a.ldy + c4 * 4);
          if (relu) {
            if (!(yv.x > 0.f)) g.x = 0.f;
            if (!(yv.y > 0.f)) g.y = 0.f;
            if (!(yv.z > 0.f)) g.z = 0.f;
            if (!(yv.w > 0.f)) g.w = 0.f;
          }
          dot = fmaf(g.x, yv.x, fmaf(g.y, yv.y, fmaf(g.z, yv.z, fmaf(g.w, yv.w, dot))));
        }
      }
      dot = warp_sum(dot);
    }
    const bool clamped = !(r > kEpsNormB);
    const float ir = clamped ? 1.f / kEpsNormB : 1.f / r;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c4 = lane + 32 * k;
      if (c4 < d4) {
        float4 g = ldg(b, n, row, c4 * 4);
        const float4 yv = ld4(a.y + row * a.ldy + c4 * 4);
        if (relu) {
          if (!(yv.x > 0.f)) g.x = 0.f;
          if (!(yv.y > 0.f)) g.y = 0.f;
          if (!(yv.z > 0.f)) g.z = 0.f;
          if (!(yv.w > 0.f)) g.w = 0.f;
        }
        if (normalize) {
          if (clamped) {
            g.x /= kEpsNormB; g.y /= kEpsNormB; g.z /= kEpsNormB; g.w /= kEpsNormB;
          } else {
            g.x = (g.x - yv.x * dot) * ir; g.y = (g.y - yv.y * dot) * ir;
            g.z = (g.z - yv.z * dot) * ir; g.w = (g.w - yv.w * dot) * ir;
          }
        }
        store_dv(a, row, c4 * 4, g);
        cs[k].x += g.x; cs[k].y += g.y; cs[k].z += g.z; cs[k].w += g.w;
      }
    }
  }
  if (a.part != nullptr) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) *reinterpret_cast<float4*>(&colacc[warp][(lane + 32 * k) * 4]) = cs[k];
    __syncthreads();
    for (int cidx = threadIdx.x; cidx < a.d; cidx += blockDim.x)
      a.part[(long long)blockIdx.x * a.d + cidx] = (colacc[0][cidx] + colacc[1][cidx]) + (colacc[2][cidx] + colacc[3][cidx]);
  }
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static bool al8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }

// shape-level eligibility for the vectorised kernels (pointer alignment is checked per call)
static bool shape_fast(int B, int d, int bn, int* CS_out) {
  if (d % 4 != 0 || d > (bn ? 512 : 2048)) return false;
  if (!bn) return true;
  const int d4 = d / 4;
  if (d4 > 32 || (d4 & (d4 - 1)) != 0) return false;     // a row must fit a power-of-two slice of one warp
  const int rstep = 256 / d4;
  static int max_vpt = 0;                                // rows per thread before the node is split over more CTAs
  if (max_vpt == 0) {
    const char* e = getenv("GP_LBWD_VPT");
    max_vpt = e != nullptr ? atoi(e) : 8;
    if (max_vpt != 1 && max_vpt != 2 && max_vpt != 4 && max_vpt != 8) max_vpt = 8;
  }
  int CS = 1;
  while (CS <= 8 && ((B + CS - 1) / CS + rstep - 1) / rstep > max_vpt) CS *= 2;
  if (CS > 8 && max_vpt < 8) {                           // cluster limit reached: fall back to more rows per thread
    CS = 1;
    while (CS <= 8 && ((B + CS - 1) / CS + rstep - 1) / rstep > 8) CS *= 2;
  }
  if (CS > 8) return false;
  if (CS_out) *CS_out = CS;
  return true;
}

// floats gp_gcn_layer_bwd_x needs in `ws` when db != NULL
static long long ws_floats(int B, int N, int d, int bn) {
  long long blocks = (long long)N * 8;                   // upper bound on CTAs of either kernel
  if (blocks < kNumSMs * 8) blocks = kNumSMs * 8;
  long long f = blocks * d + 256LL * d + (long long)N * 128;   // + batch-split partials of the generic path
  if (!shape_fast(B, d, bn, nullptr)) f += (long long)B * N * d;   // generic path: fp32 dV scratch for the column sums
  return f;
}

template <int VPT>
static int launch_bn(const LbArgs& a, int CS, int rpc, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.N * CS));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (a.h != nullptr) GP_CUDA(cudaLaunchKernelEx(&cfg, layer_bwd_bn_kernel<VPT, true>, a, CS, rpc));
  else                GP_CUDA(cudaLaunchKernelEx(&cfg, layer_bwd_bn_kernel<VPT, false>, a, CS, rpc));
  g_launches++;
  return GP_OK;
}

// full eligibility of one call for the vectorised kernels: shape AND pointer / stride alignment
static bool fast_eligible(const gp_layer_bwd* q, int* CS_out) {
  if (!al16(q->y) || q->ldy % 4 != 0) return false;
  if (q->dz != nullptr && (!(q->dz_bf16 ? al8(q->dz) : al16(q->dz)) || q->lddz % 4 != 0)) return false;
  if (q->dxn != nullptr && (!(q->dxn_bf16 ? al8(q->dxn) : al16(q->dxn)) || q->lddxn % 4 != 0)) return false;
  if (q->dout != nullptr && (!al16(q->dout) || !al16(q->argidx) || q->ldo % 4 != 0)) return false;
  if (q->h != nullptr && (!al16(q->h) || q->ldh % 4 != 0)) return false;
  if (q->dv != nullptr && !al16(q->dv)) return false;
  if (q->dv_bf16 != nullptr && ((reinterpret_cast<uintptr_t>(q->dv_bf16) & 7) != 0 || q->lddvb % 4 != 0)) return false;
  if (q->bn && q->h == nullptr && q->mean == nullptr) return false;
  return shape_fast(q->B, q->d, q->bn, CS_out);
}

int layer_bwd_fast(const gp_layer_bwd* q, cudaStream_t st, bool* handled) {
  *handled = false;
  const int d = q->d;
  int CS = 1;
  if (!fast_eligible(q, &CS)) return GP_OK;
  LbArgs a;
  a.dz = q->dz; a.lddz = q->lddz; a.dxn = q->dxn; a.lddxn = q->lddxn > 0 ? q->lddxn : (long long)q->d; a.dout = q->dout; a.argidx = q->argidx; a.ldo = q->ldo;
  a.h = q->h; a.ldh = q->ldh; a.y = q->y; a.ldy = q->ldy; a.rnorm = q->rnorm; a.mean = q->mean; a.invstd = q->invstd;
  a.B = q->B; a.N = q->N; a.d = d; a.relu = q->relu; a.bn = q->bn; a.normalize = q->normalize;
  a.dv = q->dv; a.dvb = reinterpret_cast<__nv_bfloat16*>(q->dv_bf16); a.lddvb = q->lddvb;
  a.part = q->db != nullptr ? q->ws : nullptr;
  a.nbz = q->bn ? nullptr : q->nb_zero;
  a.dz_bf16 = q->dz != nullptr && q->dz_bf16; a.dxn_bf16 = q->dxn != nullptr && q->dxn_bf16;
  long long part_rows = 0;
  static int use_cta = -1;
  if (use_cta < 0) { const char* e = getenv("GP_LBWD_CTA"); use_cta = (e == nullptr || atoi(e) != 0) ? 1 : 0; }
  if (q->bn && use_cta && q->h == nullptr && q->mean != nullptr) {
    const int d4 = d / 4;
    // bf16-only output, d >= 64 (column sums need tid < d <= 512): two 512-thread CTAs per SM with a bf16 stash
    static int use_cta2 = -1;
    if (use_cta2 < 0) { const char* e = getenv("GP_LBWD_CTA2"); use_cta2 = (e != nullptr && atoi(e) == 0) ? 0 : 1; }
    const int rstep2 = 512 / d4;
    const int vpt2 = (q->B + rstep2 - 1) / rstep2;
    if (use_cta2 && a.dv == nullptr && a.dvb != nullptr && vpt2 > 4 && vpt2 <= 16 && d <= 512) {
      static const bool no_full = getenv("GP_LBWD_NOFULL") != nullptr;
      static const bool no_cfg = getenv("GP_LBWD_NOCFG") != nullptr;
      const bool full = !no_full && q->B == (vpt2 <= 8 ? 8 : 16) * rstep2;
      // compile-time source configuration: both gradient sources present with the same element type, ReLU + normalize
      int cfg = -1;
      if (full && !no_cfg && a.dz != nullptr && a.dxn != nullptr && a.relu && a.normalize && a.dz_bf16 == a.dxn_bf16)
        cfg = (a.dz_bf16 ? 3 : 0) | (a.dout != nullptr ? 4 : 0);
#define GP_CTA2(V_) \
      do { \
        if (cfg == 0) layer_bwd_bn_cta2_kernel<V_, true, 0><<<q->N, 512, 0, st>>>(a); \
        else if (cfg == 4) layer_bwd_bn_cta2_kernel<V_, true, 4><<<q->N, 512, 0, st>>>(a); \
        else if (cfg == 3) layer_bwd_bn_cta2_kernel<V_, true, 3><<<q->N, 512, 0, st>>>(a); \
        else if (cfg == 7) layer_bwd_bn_cta2_kernel<V_, true, 7><<<q->N, 512, 0, st>>>(a); \
        else if (full) layer_bwd_bn_cta2_kernel<V_, true><<<q->N, 512, 0, st>>>(a); \
        else layer_bwd_bn_cta2_kernel<V_><<<q->N, 512, 0, st>>>(a); \
      } while (0)
      if (vpt2 <= 8) GP_CTA2(8); else GP_CTA2(16);
#undef GP_CTA2
      GP_LAUNCHED();
      part_rows = q->N;
      if (q->db != nullptr)
        GP_TRY(colsum(q->ws, part_rows, d, d, q->db, 0, q->ws + part_rows * d, st));
      *handled = true;
      return GP_OK;
    }
    const int rstep = 1024 / d4;
    const int vpt = (q->B + rstep - 1) / rstep;
    if (vpt <= 8) {
      // the gradient sources are read once: streaming loads (measured 0.325 -> 0.316 ms at cfg4; GP_LBWD_STREAM=0: off)
      static int stream_ld = -1;
      if (stream_ld < 0) { const char* e = getenv("GP_LBWD_STREAM"); stream_ld = (e != nullptr && atoi(e) == 0) ? 0 : 1; }
      if (vpt <= 2) layer_bwd_bn_cta_kernel<2><<<q->N, 1024, 0, st>>>(a);
      else if (vpt <= 4) layer_bwd_bn_cta_kernel<4><<<q->N, 1024, 0, st>>>(a);
      else if (stream_ld) layer_bwd_bn_cta_kernel<8, true><<<q->N, 1024, 0, st>>>(a);
      else layer_bwd_bn_cta_kernel<8><<<q->N, 1024, 0, st>>>(a);
      GP_LAUNCHED();
      part_rows = q->N;
      if (q->db != nullptr)
        GP_TRY(colsum(q->ws, part_rows, d, d, q->db, 0, q->ws + part_rows * d, st));
      *handled = true;
      return GP_OK;
    }
  }
  static int use_pipe = -1;
  // opt-in (GP_LBWD_PIPE=1): measured r1 on B=256, N=2048, d=128: 0.371 ms (2 stages, 512 threads) vs 0.380 ms for
  // the one-node-per-cluster kernel -- both ~2.5 TB/s; the per-node cluster sync + reduction chain bounds them,
  // not HBM.  Kept for the next step (several nodes per iteration to amortise the sync).
  if (use_pipe < 0) { const char* e = getenv("GP_LBWD_PIPE"); use_pipe = (e != nullptr && atoi(e) != 0) ? 1 : 0; }
  if (q->bn && use_pipe && q->h == nullptr && q->mean != nullptr && !a.dz_bf16 && !a.dxn_bf16) {
    // persistent pipelined variant when two stages of a CTA's rows fit in shared memory
    int PCS = 1;
    const int nstreams = 1 + (q->dz ? 1 : 0) + (q->dxn ? 1 : 0);
    auto stage_bytes = [&](int cs_) {
      const size_t rpc_ = (size_t)((q->B + cs_ - 1) / cs_);
      return (rpc_ * d * nstreams + ((rpc_ + 3) & ~(size_t)3)) * 4;
    };
    static int pipe_stages = 0;
    if (pipe_stages == 0) {
      const char* e = getenv("GP_LBWD_STAGES");
      pipe_stages = e != nullptr ? atoi(e) : 2;
      if (pipe_stages < 2 || pipe_stages > 4) pipe_stages = 2;
    }
    while (PCS <= 8 && stage_bytes(PCS) * pipe_stages > 200 * 1024) PCS *= 2;
    if (PCS <= 8 && q->N >= 64) {
      const int rpc = (q->B + PCS - 1) / PCS;
      const size_t smem = stage_bytes(PCS) * pipe_stages;
      int ncl = kNumSMs / PCS;
      if (ncl > q->N) ncl = q->N;
      static int pipe_threads = 0;
      if (pipe_threads == 0) {
        const char* e = getenv("GP_LBWD_THREADS");
        pipe_threads = e != nullptr ? atoi(e) : 512;
        if (pipe_threads != 256 && pipe_threads != 512 && pipe_threads != 1024) pipe_threads = 512;
      }
      using PipeFn = void (*)(const LbArgs, int, int, int);
      PipeFn pipe_kern = nullptr;
#define GP_PIPE_SEL(T_, S_) if (pipe_threads == T_ && pipe_stages == S_) pipe_kern = layer_bwd_bn_pipe_kernel<T_, S_>;
      GP_PIPE_SEL(256, 2) GP_PIPE_SEL(256, 3) GP_PIPE_SEL(256, 4)
      GP_PIPE_SEL(512, 2) GP_PIPE_SEL(512, 3) GP_PIPE_SEL(512, 4)
      GP_PIPE_SEL(1024, 2) GP_PIPE_SEL(1024, 3) GP_PIPE_SEL(1024, 4)
#undef GP_PIPE_SEL
      GP_CUDA(cudaFuncSetAttribute(pipe_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));   // one site, several variants
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(ncl * PCS));
      cfg.blockDim = dim3(pipe_threads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = PCS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      // persistent clusters must all be co-resident: GPC sizes are not multiples of every cluster size
      static int max_clusters[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      if (max_clusters[PCS] == 0) {
        int mc = 0;
        if (cudaOccupancyMaxActiveClusters(&mc, pipe_kern, &cfg) != cudaSuccess || mc < 1) mc = 1;
        max_clusters[PCS] = mc;
        if (getenv("GP_DEBUG")) fprintf(stderr, "[gp] layer_bwd_bn_pipe: cluster %d, smem %zu B -> max active clusters %d\n", PCS, smem, mc);
      }
      if (ncl > max_clusters[PCS]) ncl = max_clusters[PCS];
      cfg.gridDim = dim3((unsigned)(ncl * PCS));
      GP_CUDA(cudaLaunchKernelEx(&cfg, pipe_kern, a, PCS, rpc, ncl));
      g_launches++;
      part_rows = (long long)ncl * PCS;
      if (q->db != nullptr)
        GP_TRY(colsum(q->ws, part_rows, d, d, q->db, 0, q->ws + part_rows * d, st));
      *handled = true;
      return GP_OK;
    }
  }
  if (q->bn) {
    const int rstep = 256 / (d / 4);
    const int rpc = (q->B + CS - 1) / CS;
    const int vpt = (rpc + rstep - 1) / rstep;
    if (vpt <= 1) GP_TRY(launch_bn<1>(a, CS, rpc, st));
    else if (vpt <= 2) GP_TRY(launch_bn<2>(a, CS, rpc, st));
    else if (vpt <= 4) GP_TRY(launch_bn<4>(a, CS, rpc, st));
    else GP_TRY(launch_bn<8>(a, CS, rpc, st));
    part_rows = (long long)q->N * CS;
  } else {
    const long long rows = (long long)q->B * q->N;
    long long blocks = (rows + 7) / 8;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    const int d4 = d / 4;
    // four 256-thread blocks per SM (64 registers) is the measured optimum for every row width: wider rows are capped
    // there with launch bounds (d = 256: 0.329 -> 0.262 ms; d = 512: 0.741 -> 0.535 ms at [B*N = 524288] rows)
    // compile-time source configuration (see the kernel): dz only (+ readout scatter), no ReLU, normalize, full-width rows
    static const bool no_cfg_row = getenv("GP_LBWD_NOCFG") != nullptr;
    int rcfg = -1;
    if (!no_cfg_row && a.dz != nullptr && a.dxn == nullptr && !a.relu && a.normalize && a.dv == nullptr && (d4 == 32 || d4 == 128))
      rcfg = (a.dz_bf16 ? 1 : 0) | (a.dout != nullptr ? 4 : 0);
    if (rcfg >= 0 && d4 == 32) {
      if (rcfg == 0) layer_bwd_row_kernel<1, 1, 0><<<(int)blocks, 256, 0, st>>>(a, rows);
      else if (rcfg == 1) layer_bwd_row_kernel<1, 1, 1><<<(int)blocks, 256, 0, st>>>(a, rows);
      else if (rcfg == 4) layer_bwd_row_kernel<1, 1, 4><<<(int)blocks, 256, 0, st>>>(a, rows);
      else layer_bwd_row_kernel<1, 1, 5><<<(int)blocks, 256, 0, st>>>(a, rows);
    } else if (rcfg >= 0 && d4 == 128) {
      if (rcfg == 0) layer_bwd_row_kernel<4, 4, 0><<<(int)blocks, 256, 0, st>>>(a, rows);
      else if (rcfg == 1) layer_bwd_row_kernel<4, 4, 1><<<(int)blocks, 256, 0, st>>>(a, rows);
      else if (rcfg == 4) layer_bwd_row_kernel<4, 4, 4><<<(int)blocks, 256, 0, st>>>(a, rows);
      else layer_bwd_row_kernel<4, 4, 5><<<(int)blocks, 256, 0, st>>>(a, rows);
    } else
    if (d4 <= 32) layer_bwd_row_kernel<1><<<(int)blocks, 256, 0, st>>>(a, rows);
    else if (d4 <= 64) layer_bwd_row_kernel<2, 4><<<(int)blocks, 256, 0, st>>>(a, rows);
    else if (d4 <= 128) {
      // 16 floats x 2 operands per lane: left alone the compiler takes 106 registers (2 blocks per SM, 0.74 ms at
      // [B*N, 512]); capped at 64 registers four blocks are resident and the pass runs at 5.0 TB/s (0.535 ms).
      // Measured: 3 blocks 0.66 ms, 5 blocks (48 registers, spills) 0.64 ms, 6 blocks 0.75 ms.
      static int minb = -1;
      if (minb < 0) { const char* e = getenv("GP_LBWD_ROW_MINB"); minb = e != nullptr ? atoi(e) : 4; }
      if (minb >= 4) layer_bwd_row_kernel<4, 4><<<(int)blocks, 256, 0, st>>>(a, rows);
      else layer_bwd_row_kernel<4><<<(int)blocks, 256, 0, st>>>(a, rows);
    }
    else if (d4 <= 256) layer_bwd_row_wide_kernel<8><<<(int)blocks, 128, 0, st>>>(a, rows);
    else if (!no_cfg_row && a.dz != nullptr && a.dxn == nullptr && a.dout == nullptr && !a.relu && a.normalize) {
      if (a.dz_bf16) layer_bwd_row_wide_kernel<16, 6, 1><<<(int)blocks, 128, 0, st>>>(a, rows);
      else layer_bwd_row_wide_kernel<16, 6, 0><<<(int)blocks, 128, 0, st>>>(a, rows);
    }
    else {
      static int wminb = -1;
      if (wminb < 0) { const char* e = getenv("GP_WIDE_MINB"); wminb = e != nullptr ? atoi(e) : 1; }   /* 80 registers: six / three blocks per SM, measured 2.31 -> 1.95 ms at cfg5 */
      if (wminb) layer_bwd_row_wide_kernel<16, 6><<<(int)blocks, 128, 0, st>>>(a, rows);
      else layer_bwd_row_wide_kernel<16><<<(int)blocks, 128, 0, st>>>(a, rows);
    }
    GP_LAUNCHED();
    part_rows = blocks;
  }
  if (q->db != nullptr)
    GP_TRY(colsum(q->ws, part_rows, d, d, q->db, 0, q->ws + part_rows * d, st));
  *handled = true;
  return GP_OK;
}

}  // namespace gp

using namespace gp;

extern "C" long long gp_gcn_layer_bwd_ws(int B, int N, int d, int bn) { return ws_floats(B, N, d, bn); }
// 1 if a layer of this shape runs on the vectorised kernels (given aligned operands): the only ones that take bf16 dz / dxn
extern "C" int gp_gcn_layer_bwd_vectorised(int B, int d, int bn) { return shape_fast(B, d, bn, nullptr) ? 1 : 0; }
// 1 if a layer of this shape is served by a kernel instantiated for bf16 gradient sources (compile-time source
// configuration): BatchNorm layers on the two-CTAs-per-SM kernel with a batch that fills its passes exactly, last layers
// with rows of 128 or 512 floats.  Elsewhere bf16 sources run through run-time branches and are SLOWER than fp32 ones
// (measured: 1256-wide rows 2.8 vs 1.9 ms), so callers keep fp32 intermediates there.
extern "C" int gp_gcn_layer_bwd_bf16_sources_fast(int B, int d, int bn) {
  if (!shape_fast(B, d, bn, nullptr)) return 0;
  if (!bn) return (d == 128 || d == 512) ? 1 : 0;
  if (d > 512) return 0;
  const int rstep2 = 512 / (d / 4);
  return (B == 8 * rstep2 || B == 16 * rstep2) ? 1 : 0;
}

// exact workspace of ONE call (ws / db fields ignored): the generic kernel borrows an fp32 dV from ws when the
// caller asked for the bias gradient without an fp32 dV and the operands miss the vectorised path's alignment.
extern "C" long long gp_gcn_layer_bwd_ws_x(const gp_layer_bwd* q) {
  if (q == nullptr) return 0;
  long long f = ws_floats(q->B, q->N, q->d, q->bn);
  if (shape_fast(q->B, q->d, q->bn, nullptr) && !fast_eligible(q, nullptr) && q->dv == nullptr)
    f += (long long)q->B * q->N * q->d;
  return f;
}

extern "C" int gp_gcn_layer_bwd_x(const gp_layer_bwd* q, gp_stream_t stream) {
  GP_REQUIRE(q != nullptr, "gcn_layer_bwd_x: null descriptor");
  GP_REQUIRE(q->y && (q->dv || q->dv_bf16) && q->B > 0 && q->N > 0 && q->d > 0, "gcn_layer_bwd_x: bad args");
  GP_REQUIRE((long long)q->B * q->N <= 0x7fffffffLL, "gcn_layer_bwd_x: B * N must stay below 2^31 rows");
  GP_REQUIRE(!q->bn || (q->invstd && (q->h || q->mean)), "gcn_layer_bwd_x: bn needs invstd and h or mean");
  GP_REQUIRE(!q->normalize || q->rnorm, "gcn_layer_bwd_x: normalize needs rnorm");
  GP_REQUIRE(!q->dout || q->argidx, "gcn_layer_bwd_x: dout needs argidx");
  GP_REQUIRE(!q->db || q->ws, "gcn_layer_bwd_x: db needs ws (gp_gcn_layer_bwd_ws_x floats)");
  bool handled = false;
  GP_TRY(layer_bwd_fast(q, S(stream), &handled));
  if (handled) return GP_OK;
  GP_REQUIRE(!q->dz_bf16 && !q->dxn_bf16, "gcn_layer_bwd_x: bf16 gradient sources need the vectorised path (d %% 4 == 0, aligned rows)");
  // generic path: with db but no fp32 dV it borrows B*N*d floats from ws -- size ws with gp_gcn_layer_bwd_ws_x
  return layer_bwd_generic(q, S(stream));
}
