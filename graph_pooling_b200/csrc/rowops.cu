// Row / node-index reductions and epilogues of the DiffPool path (fp32, HBM-bound):
//   bias + L2 normalize (encoders.py:323-326), ReLU + BatchNorm-per-node-index with batch
//   statistics (encoders.py:1062-1064,1048-1052) and its backward fused with the ReLU mask, the
//   max-readout scatter and the normalize backward, max readout (encoders.py:1097,1257,1287),
//   masked assignment softmax (encoders.py:1273-1275), cross entropy (encoders.py:1127), colsum.
// All reductions are warp-shuffle based over coalesced rows; no atomics, deterministic.
#include <cuda_bf16.h>
#include "common.cuh"

namespace gp {

constexpr float kEpsNorm = 1e-12f;
constexpr float kEpsBn = 1e-5f;
constexpr size_t kNodeCacheMax = 200 * 1024;   // per-node working set staged in shared memory (<= 227 KB/CTA)

// ---------------------------------------------------------------------------------------------
// V (+bias) -> Y = V / max(||V||, eps), in place; one warp per row.
// ---------------------------------------------------------------------------------------------
__global__ void bias_normalize_kernel(float* __restrict__ v, const float* __restrict__ bias,
                                      float* __restrict__ rnorm, long long rows, int d, long long ld,
                                      int normalize) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    float* p = v + r * ld;
    float ss = 0.f;
    for (int c = lane; c < d; c += 32) {
      float x = p[c];
      if (bias != nullptr) { x += bias[c]; p[c] = x; }
      ss = fmaf(x, x, ss);
    }
    if (!normalize) { if (lane == 0 && rnorm) rnorm[r] = 1.f; continue; }
    ss = warp_sum(ss);
    const float nrm = fmaxf(sqrtf(ss), kEpsNorm);
    if (lane == 0 && rnorm) rnorm[r] = nrm;
    for (int c = lane; c < d; c += 32) p[c] = p[c] / nrm;
  }
}

// ---------------------------------------------------------------------------------------------
// ReLU + BN over (batch, feature) per node index n.  One block per node (grid.x = N).
// Two-pass mean / variance (the block re-reads its B*d elements from L1/L2), third pass writes.
// ---------------------------------------------------------------------------------------------
// CACHE: the node's B*d activations are staged once in dynamic shared memory, so HBM is read once.
template <bool CACHE>
__global__ void relu_bn_fwd_kernel(const float* __restrict__ y, float* __restrict__ h, long long ldh,
                                   float* __restrict__ mean, float* __restrict__ invstd,
                                   int B, int N, int d, int relu, int bn) {
  extern __shared__ float cache[];
  __shared__ float sh[33];
  const int n = blockIdx.x;
  const int total = B * d;
  float mu = 0.f, is = 1.f;
  if (bn) {
    float s = 0.f;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int b = i / d, c = i - b * d;
      float x = y[((long long)b * N + n) * d + c];
      if (relu) x = fmaxf(x, 0.f);
      if (CACHE) cache[i] = x;
      s += x;
    }
    mu = block_sum(s, sh) / (float)total;
    float q = 0.f;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      float x;
      if (CACHE) {
        x = cache[i];
      } else {
        const int b = i / d, c = i - b * d;
        x = y[((long long)b * N + n) * d + c];
        if (relu) x = fmaxf(x, 0.f);
      }
      const float t = x - mu;
      q = fmaf(t, t, q);
    }
    const float var = block_sum(q, sh) / (float)total;
    is = 1.0f / sqrtf(var + kEpsBn);
    if (threadIdx.x == 0) { mean[n] = mu; invstd[n] = is; }
  }
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int b = i / d, c = i - b * d;
    float x;
    if (CACHE && bn) {
      x = cache[i];
    } else {
      x = y[((long long)b * N + n) * d + c];
      if (relu) x = fmaxf(x, 0.f);
    }
    h[((long long)b * N + n) * ldh + c] = (x - mu) * is;
  }
}

// ---------------------------------------------------------------------------------------------
// Batch-split variants for SMALL N and LARGE B (ENZYMES-sized graphs in big batches): one CTA per node index
// would leave most of the 148 SMs idle (N = 100 -> 100 CTAs, each walking B*d elements), so a node's batch is
// cut into S slices: grid (N, S).  Statistics are combined with Chan's parallel mean / M2 update (exact
// two-pass quality, no E[x^2] - mean^2 cancellation); everything stays deterministic.
// ---------------------------------------------------------------------------------------------
__global__ void bn_split_stats_kernel(const float* __restrict__ y, int B, int N, int d, int relu, int bs,
                                      float* __restrict__ part) {
  __shared__ float sh[33];
  const int n = blockIdx.x, sidx = blockIdx.y;
  const int b0 = sidx * bs, b1 = min(B, b0 + bs);
  const int total = (b1 - b0) * d;
  float s = 0.f;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int b = b0 + i / d, c = i % d;
    float x = y[((long long)b * N + n) * d + c];
    if (relu) x = fmaxf(x, 0.f);
    s += x;
  }
  const float mu = total > 0 ? block_sum(s, sh) / (float)total : 0.f;
  float q = 0.f;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int b = b0 + i / d, c = i % d;
    float x = y[((long long)b * N + n) * d + c];
    if (relu) x = fmaxf(x, 0.f);
    const float t = x - mu;
    q = fmaf(t, t, q);
  }
  const float m2 = block_sum(q, sh);
  if (threadIdx.x == 0) {
    float* o = part + ((long long)n * gridDim.y + sidx) * 3;
    o[0] = (float)total; o[1] = mu; o[2] = m2;
  }
}

__global__ void bn_split_apply_kernel(const float* __restrict__ y, float* __restrict__ h, long long ldh,
                                      float* __restrict__ mean, float* __restrict__ invstd, int B, int N, int d,
                                      int relu, int bs, const float* __restrict__ part) {
  __shared__ float sh_mu, sh_is;
  const int n = blockIdx.x, sidx = blockIdx.y, S = gridDim.y;
  if (threadIdx.x == 0) {                                // Chan et al. combination, same order in every CTA
    double cnt = 0.0, mu = 0.0, m2 = 0.0;
    for (int j = 0; j < S; ++j) {
      const float* o = part + ((long long)n * S + j) * 3;
      const double c = o[0], m = o[1], q = o[2];
      if (c <= 0.0) continue;
      const double tot = cnt + c, dlt = m - mu;
      m2 += q + dlt * dlt * cnt * c / tot;
      mu += dlt * c / tot;
      cnt = tot;
    }
    const float is = (float)(1.0 / sqrt(m2 / cnt + (double)kEpsBn));
    sh_mu = (float)mu; sh_is = is;
    if (sidx == 0) { mean[n] = (float)mu; invstd[n] = is; }
  }
  __syncthreads();
  const float mu = sh_mu, is = sh_is;
  const int b0 = sidx * bs, b1 = min(B, b0 + bs);
  const int total = (b1 - b0) * d;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int b = b0 + i / d, c = i % d;
    float x = y[((long long)b * N + n) * d + c];
    if (relu) x = fmaxf(x, 0.f);
    h[((long long)b * N + n) * ldh + c] = (x - mu) * is;
  }
}

// number of batch slices for the split kernels (0 = keep the one-CTA-per-node kernels)
static int split_slices(int B, int N, int d) {
  const long long total = (long long)B * d;
  if (N >= 2 * kNumSMs || total < 16384) return 0;
  long long S = (total + 4095) / 4096;                   // ~4K elements per CTA
  const long long want = (4LL * kNumSMs + N - 1) / N;    // enough CTAs for ~4 per SM
  if (S > want) S = want;
  if (S > 64) S = 64;
  if (S > B) S = B;
  return S < 2 ? 0 : (int)S;
}

// ---------------------------------------------------------------------------------------------
// Backward of [slot + readout scatter + next-layer dX] -> BN -> ReLU -> normalize.
// One block per node index n; pass 1 block-reduces mean(g) and mean(g*Hhat); pass 2 is one warp
// per (b, n) row: dR, ReLU mask, <Y,dY> by shuffle, dV.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float layer_g(const float* dz, long long lddz, const float* dxn, long long lddxn,
                                         const float* dout, const int32_t* argidx, long long ldo, int b, int n, int c,
                                         int N, int d) {
  float g = 0.f;
  const long long row = (long long)b * N + n;
  if (dz != nullptr) g += dz[row * lddz + c];
  if (dxn != nullptr) g += dxn[row * lddxn + c];
  if (dout != nullptr && argidx[(long long)b * ldo + c] == n) g += dout[(long long)b * ldo + c];
  return g;
}

// CACHE: pass 1 stages the combined gradient g of the node's B*d entries in dynamic shared memory, so the
// gradient sources are read from HBM once.  MAXE > 0: a row (d <= 32*MAXE) stays in registers between the
// <Y,dY> reduction and the final write, so h / y are read once in pass 2.
// Hhat of one element: the stored BN output if available, else recomputed from Y and the saved statistics.
__device__ __forceinline__ float hhat_of(const float* h, long long ldh, const float* y, long long ldy, long long row,
                                         int c, int relu, float mu, float is) {
  if (h != nullptr) return h[row * ldh + c];
  float x = y[row * ldy + c];
  if (relu) x = fmaxf(x, 0.f);
  return (x - mu) * is;
}
__device__ __forceinline__ void put_dv(float* dv, __nv_bfloat16* dvb, long long lddvb, long long row, int d, int c,
                                       float g) {
  if (dv != nullptr) dv[row * d + c] = g;
  if (dvb != nullptr) dvb[row * lddvb + c] = __float2bfloat16_rn(g);
}

struct LbG {
  const float* dz; long long lddz; const float* dxn; long long lddxn;
  const float* dout; const int32_t* argidx; long long ldo;
  const float* h; long long ldh; const float* y; long long ldy;
  const float* rnorm; const float* mean; const float* invstd;
  int B, N, d, relu, bn, normalize;
  float* dv; __nv_bfloat16* dvb; long long lddvb;
};

// rows b_lo..b_hi-1 of node n: dR (BN backward with the batch means m1, m2), ReLU mask, <Y,dY> by shuffle, dV.
// One warp per row.  cache: optional per-node staging of g (indexed [b * d + c]) or nullptr.
template <int MAXE>
__device__ __forceinline__ void layer_bwd_rows(const LbG& a, int n, int b_lo, int b_hi, float m1, float m2, float mu,
                                               float is, const float* cache) {
  const int d = a.d, N = a.N;
  const bool cached = cache != nullptr;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int b = b_lo + w; b < b_hi; b += nw) {
    const long long row = (long long)b * N + n;
    float r = 1.f;
    bool clamped = false;
    if (a.normalize) {
      r = a.rnorm[row];
      clamped = !(r > kEpsNorm);
    }
    if (MAXE > 0) {
      float gv[MAXE > 0 ? MAXE : 1], yv[MAXE > 0 ? MAXE : 1];
      float dot = 0.f;
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        const int c = lane + 32 * e;
        float g = 0.f, yy = 0.f;
        if (c < d) {
          g = cached ? cache[b * d + c] : layer_g(a.dz, a.lddz, a.dxn, a.lddxn, a.dout, a.argidx, a.ldo, b, n, c, N, d);
          if (a.bn) g = (g - m1 - hhat_of(a.h, a.ldh, a.y, a.ldy, row, c, a.relu, mu, is) * m2) * is;
          yy = a.y[row * a.ldy + c];
          if (a.relu && !(yy > 0.f)) g = 0.f;
          dot = fmaf(g, yy, dot);
        }
        gv[e] = g; yv[e] = yy;
      }
      if (a.normalize) dot = warp_sum(dot);
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        const int c = lane + 32 * e;
        if (c < d) {
          float g = gv[e];
          if (a.normalize) g = clamped ? g / kEpsNorm : (g - yv[e] * dot) / r;
          put_dv(a.dv, a.dvb, a.lddvb, row, d, c, g);
        }
      }
    } else {
      float dot = 0.f;
      for (int c = lane; c < d; c += 32) {
        float g = cached ? cache[b * d + c] : layer_g(a.dz, a.lddz, a.dxn, a.lddxn, a.dout, a.argidx, a.ldo, b, n, c, N, d);
        if (a.bn) g = (g - m1 - hhat_of(a.h, a.ldh, a.y, a.ldy, row, c, a.relu, mu, is) * m2) * is;
        const float yy = a.y[row * a.ldy + c];
        if (a.relu && !(yy > 0.f)) g = 0.f;
        dot = fmaf(g, yy, dot);
      }
      if (a.normalize) dot = warp_sum(dot);
      for (int c = lane; c < d; c += 32) {
        float g = cached ? cache[b * d + c] : layer_g(a.dz, a.lddz, a.dxn, a.lddxn, a.dout, a.argidx, a.ldo, b, n, c, N, d);
        if (a.bn) g = (g - m1 - hhat_of(a.h, a.ldh, a.y, a.ldy, row, c, a.relu, mu, is) * m2) * is;
        const float yy = a.y[row * a.ldy + c];
        if (a.relu && !(yy > 0.f)) g = 0.f;
        if (a.normalize) g = clamped ? g / kEpsNorm : (g - yy * dot) / r;
        put_dv(a.dv, a.dvb, a.lddvb, row, d, c, g);
      }
    }
  }
}

template <bool CACHE, int MAXE>
__global__ void gcn_layer_bwd_kernel(const LbG a) {
  extern __shared__ float cache[];
  __shared__ float sh[33];
  const int n = blockIdx.x;
  const int B = a.B, N = a.N, d = a.d;
  const int total = B * d;
  float m1 = 0.f, m2 = 0.f, is = 1.f, mu = 0.f;
  if (a.bn) {
    is = a.invstd[n];
    if (a.mean != nullptr) mu = a.mean[n];
    float s1 = 0.f, s2 = 0.f;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int b = i / d, c = i - b * d;
      const float g = layer_g(a.dz, a.lddz, a.dxn, a.lddxn, a.dout, a.argidx, a.ldo, b, n, c, N, d);
      if (CACHE) cache[i] = g;
      s1 += g;
      s2 = fmaf(g, hhat_of(a.h, a.ldh, a.y, a.ldy, (long long)b * N + n, c, a.relu, mu, is), s2);
    }
    m1 = block_sum(s1, sh) / (float)total;
    m2 = block_sum(s2, sh) / (float)total;
  }
  layer_bwd_rows<MAXE>(a, n, 0, B, m1, m2, mu, is, (CACHE && a.bn) ? cache : nullptr);
}

// batch-split variants (see bn_split_stats_kernel): grid (N, S); partial sums of g and g*Hhat per slice, then the
// row pass with the combined batch means (slices summed in a fixed order: deterministic).
__global__ void layer_bwd_split_stats_kernel(const LbG a, int bs, float* __restrict__ part) {
  __shared__ float sh[33];
  const int n = blockIdx.x, sidx = blockIdx.y;
  const int d = a.d, N = a.N;
  const int b0 = sidx * bs, b1 = min(a.B, b0 + bs);
  const int total = (b1 - b0) * d;
  const float is = a.invstd[n];
  const float mu = a.mean != nullptr ? a.mean[n] : 0.f;
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int b = b0 + i / d, c = i % d;
    const float g = layer_g(a.dz, a.lddz, a.dxn, a.lddxn, a.dout, a.argidx, a.ldo, b, n, c, N, d);
    s1 += g;
    s2 = fmaf(g, hhat_of(a.h, a.ldh, a.y, a.ldy, (long long)b * N + n, c, a.relu, mu, is), s2);
  }
  s1 = block_sum(s1, sh);
  s2 = block_sum(s2, sh);
  if (threadIdx.x == 0) {
    float* o = part + ((long long)n * gridDim.y + sidx) * 2;
    o[0] = s1; o[1] = s2;
  }
}

template <int MAXE>
__global__ void layer_bwd_split_apply_kernel(const LbG a, int bs, const float* __restrict__ part) {
  const int n = blockIdx.x, sidx = blockIdx.y, S = gridDim.y;
  float m1 = 0.f, m2 = 0.f, is = 1.f, mu = 0.f;
  if (a.bn) {
    is = a.invstd[n];
    if (a.mean != nullptr) mu = a.mean[n];
    double t1 = 0.0, t2 = 0.0;
    for (int j = 0; j < S; ++j) {                        // every thread: S <= 64 L1-resident values, fixed order
      t1 += (double)part[((long long)n * S + j) * 2];
      t2 += (double)part[((long long)n * S + j) * 2 + 1];
    }
    const double inv = 1.0 / ((double)a.B * (double)a.d);
    m1 = (float)(t1 * inv); m2 = (float)(t2 * inv);
  }
  const int b0 = sidx * bs, b1 = min(a.B, b0 + bs);
  layer_bwd_rows<MAXE>(a, n, b0, b1, m1, m2, mu, is, nullptr);
}

// ---------------------------------------------------------------------------------------------
// Max readout over nodes: block = 32 feature lanes x 8 row groups.
// ---------------------------------------------------------------------------------------------
// N = rows scanned per graph, pitch = rows between consecutive graphs in memory (>= N; rows [N, pitch) do not exist
// for the readout: dead clusters of a padded assignment width, gp_readout_max_fwd_x).
__global__ void readout_max_kernel(const float* __restrict__ z, long long ldz, const int32_t* __restrict__ nb,
                                   int N, int pitch, int F, float* __restrict__ out, int32_t* __restrict__ argidx,
                                   long long ldo) {
  __shared__ float sv[8][33];
  __shared__ int si[8][33];
  const int b = blockIdx.y;
  const int f = blockIdx.x * 32 + threadIdx.x;
  const int nreal = nb != nullptr ? min(nb[b], N) : N;
  float best = -INFINITY;
  int bi = -1;
  if (f < F) {
    const float* p = z + (long long)b * pitch * ldz + f;
    for (int n = threadIdx.y; n < nreal; n += 8) {
      const float v = p[(long long)n * ldz];
      if (v > best) { best = v; bi = n; }      // strict: lowest index kept within this thread
    }
  }
  sv[threadIdx.y][threadIdx.x] = best;
  si[threadIdx.y][threadIdx.x] = bi;
  __syncthreads();
  if (threadIdx.y == 0 && f < F) {
    for (int g = 1; g < 8; ++g) {
      const float v = sv[g][threadIdx.x];
      const int i = si[g][threadIdx.x];
      if (i >= 0 && (v > best || (v == best && i < bi) || bi < 0)) { best = v; bi = i; }
    }
    if (nreal < N) {                 // masked pad rows count as 0 (index nreal > every real index)
      if (bi < 0 || 0.f > best) { best = 0.f; bi = -1; }
    }
    out[(long long)b * ldo + f] = best;
    argidx[(long long)b * ldo + f] = bi;
  }
}

// ---------------------------------------------------------------------------------------------
// Masked row softmax (in place) and its backward; one warp per row.
// ---------------------------------------------------------------------------------------------
__global__ void softmax_mask_fwd_kernel(float* __restrict__ t, const int32_t* __restrict__ nb, long long rows,
                                        int N, int K) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    float* p = t + r * K;
    const int b = (int)(r / N), n = (int)(r % N);
    if (nb != nullptr && n >= nb[b]) {
      for (int c = lane; c < K; c += 32) p[c] = 0.f;
      continue;
    }
    float mx = -INFINITY;
    for (int c = lane; c < K; c += 32) mx = fmaxf(mx, p[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < K; c += 32) { const float e = expf(p[c] - mx); p[c] = e; s += e; }
    s = warp_sum(s);
    for (int c = lane; c < K; c += 32) p[c] = p[c] / s;
  }
}

__global__ void softmax_mask_bwd_kernel(const float* __restrict__ s, const float* __restrict__ ds,
                                        const int32_t* __restrict__ nb, long long rows, int N, int K,
                                        float* __restrict__ dt) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const int b = (int)(r / N), n = (int)(r % N);
    const float* sp = s + r * K;
    const float* gp = ds + r * K;
    float* o = dt + r * K;
    if (nb != nullptr && n >= nb[b]) {
      for (int c = lane; c < K; c += 32) o[c] = 0.f;
      continue;
    }
    float dot = 0.f;
    for (int c = lane; c < K; c += 32) dot = fmaf(sp[c], gp[c], dot);
    dot = warp_sum(dot);
    for (int c = lane; c < K; c += 32) o[c] = sp[c] * (gp[c] - dot);
  }
}

// ---------------------------------------------------------------------------------------------
// Cross entropy (mean over batch), one block; probabilities saved for the backward.
// ---------------------------------------------------------------------------------------------
__global__ void ce_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ label, int B, int C,
                              float* __restrict__ loss, float* __restrict__ probs) {
  __shared__ float sh[33];
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float* p = logits + (long long)b * C;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, p[c]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(p[c] - mx);
    const float lse = mx + logf(s);
    for (int c = 0; c < C; ++c) probs[(long long)b * C + c] = expf(p[c] - lse);
    acc += lse - p[label[b]];
  }
  const float tot = block_sum(acc, sh);
  if (threadIdx.x == 0) *loss = tot / (float)B;
}

__global__ void ce_bwd_kernel(const float* __restrict__ probs, const int64_t* __restrict__ label,
                              const float* __restrict__ upstream, int B, int C, float* __restrict__ dlogits) {
  const float g = (upstream != nullptr ? *upstream : 1.f) / (float)B;
  const int total = B * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / C, c = i - b * C;
    dlogits[i] = g * (probs[i] - (label[b] == c ? 1.f : 0.f));
  }
}

// ---------------------------------------------------------------------------------------------
// Column sum, two deterministic stages: partial[R][d] then out[d].
// ---------------------------------------------------------------------------------------------
__global__ void colsum_stage1(const float* __restrict__ x, long long rows, int d, long long ld, int R,
                              float* __restrict__ partial) {
  __shared__ float sh[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int r = blockIdx.y;
  const long long per = (rows + R - 1) / R;
  const long long r0 = r * per, r1 = min(rows, r0 + per);
  float s = 0.f;
  if (c < d)
    for (long long i = r0 + threadIdx.y; i < r1; i += 8) s += x[i * ld + c];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < d) {
    for (int g = 1; g < 8; ++g) s += sh[g][threadIdx.x];
    partial[(long long)r * d + c] = s;
  }
}
__global__ void colsum_stage2(const float* __restrict__ partial, int R, int d, float* __restrict__ out,
                              int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  float s = 0.f;
  for (int r = 0; r < R; ++r) s += partial[(long long)r * d + c];
  out[c] = accumulate ? out[c] + s : s;
}

__global__ void relu_mask_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, long long n,
                                     float* __restrict__ dx) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}

__global__ void fill_kernel(float* __restrict__ x, long long n, float v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = v;
}
__global__ void fill_i32_kernel(int32_t* __restrict__ x, long long n, int32_t v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = v;
}
// dst [rows_dst, cols_dst] (row stride ld_dst) = src [rows, cols] (row stride ld_src) in the top-left corner, `fill`
// elsewhere
__global__ void pad_copy_kernel(const float* __restrict__ src, long long ld_src, long long rows, int cols,
                                float* __restrict__ dst, long long ld_dst, long long rows_dst, int cols_dst,
                                float fill) {
  const long long total = rows_dst * cols_dst;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols_dst;
    const int c = (int)(i - r * cols_dst);
    dst[r * ld_dst + c] = (r < rows && c < cols) ? src[r * ld_src + c] : fill;
  }
}
// Dropout of a GraphConv input (encoders.py:316-317, nn.Dropout in training mode): y = keep ? x / (1 - p) : 0 with
// keep decided by a counter-based hash of (seed, row * d + column) -- the same call with the same seed reproduces
// the mask, which is how the backward applies it to the input gradient (in place: y == x).
__device__ __forceinline__ bool dropout_keep(unsigned long long seed, unsigned long long idx, float p) {
  unsigned long long z = seed + (idx + 1ull) * 0x9E3779B97F4A7C15ull;              // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(unsigned)(z >> 40) * (1.f / 16777216.f) >= p;                   // 24 uniform bits
}
__global__ void dropout_kernel(const float* __restrict__ x, long long ldx, long long rows, int d, float p,
                               unsigned long long seed, float* __restrict__ y, long long ldy,
                               __nv_bfloat16* __restrict__ yb, long long ldyb) {
  const float scale = 1.f / (1.f - p);
  const long long total = rows * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / d;
    const int c = (int)(i - r * d);
    const float v = dropout_keep(seed, (unsigned long long)i, p) ? x[r * ldx + c] * scale : 0.f;
    if (y != nullptr) y[r * ldy + c] = v;
    if (yb != nullptr) yb[r * ldyb + c] = __float2bfloat16_rn(v);
  }
}
__global__ void axpy_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float a) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = fmaf(a, x[i], y[i]);
}

static inline int warp_grid(long long rows, int threads) {
  const long long per = threads / 32;
  long long blocks = (rows + per - 1) / per;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

int bias_normalize(float* v, const float* bias, float* rnorm, long long rows, int d, long long ld,
                   int normalize, cudaStream_t st) {
  bias_normalize_kernel<<<warp_grid(rows, 256), 256, 0, st>>>(v, bias, rnorm, rows, d, ld, normalize);
  GP_LAUNCHED();
  return GP_OK;
}

// few rows (per-graph partials of an ENZYMES-sized batch, ...): one pass, a thread per column
__global__ void colsum_small_kernel(const float* __restrict__ x, int rows, int d, long long ld, float* __restrict__ out,
                                    int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += x[(long long)r * ld + c];
  out[c] = accumulate ? out[c] + s : s;
}

int colsum(const float* x, long long rows, int d, long long ld, float* out, int accumulate, float* ws,
           cudaStream_t st) {
  GP_REQUIRE(x && out && ws && d > 0, "colsum: bad args");
  if (rows <= 64) {
    colsum_small_kernel<<<(d + 127) / 128, 128, 0, st>>>(x, (int)rows, d, ld, out, accumulate);
    GP_LAUNCHED();
    return GP_OK;
  }
  int R = (int)((rows + 63) / 64);
  if (R > 256) R = 256;
  if (R < 1) R = 1;
  dim3 grid((d + 31) / 32, R), block(32, 8);
  colsum_stage1<<<grid, block, 0, st>>>(x, rows, d, ld, R, ws);
  GP_LAUNCHED();
  colsum_stage2<<<(d + 127) / 128, 128, 0, st>>>(ws, R, d, out, accumulate);
  GP_LAUNCHED();
  return GP_OK;
}

}  // namespace gp

using namespace gp;

extern "C" long long gp_relu_bn_fwd_ws(int B, int N, int d) {
  return (long long)N * split_slices(B, N, d) * 3;
}

extern "C" int gp_relu_bn_fwd_x(const float* y, float* h, long long ldh, float* mean, float* invstd, int B, int N,
                                int d, int relu, int bn, float* ws, gp_stream_t stream) {
  GP_REQUIRE(y && h && B > 0 && N > 0 && d > 0 && ldh >= d, "relu_bn_fwd: bad args");
  GP_REQUIRE(!bn || (mean && invstd), "relu_bn_fwd: bn needs mean/invstd");
  const int S = (bn && ws != nullptr) ? split_slices(B, N, d) : 0;
  if (S > 0) {
    const int bs = (B + S - 1) / S;
    dim3 grid(N, S);
    bn_split_stats_kernel<<<grid, 256, 0, gp::S(stream)>>>(y, B, N, d, relu, bs, ws);
    GP_LAUNCHED();
    bn_split_apply_kernel<<<grid, 256, 0, gp::S(stream)>>>(y, h, ldh, mean, invstd, B, N, d, relu, bs, ws);
    GP_LAUNCHED();
    return GP_OK;
  }
  return gp_relu_bn_fwd(y, h, ldh, mean, invstd, B, N, d, relu, bn, stream);
}

extern "C" int gp_relu_bn_fwd(const float* y, float* h, long long ldh, float* mean, float* invstd, int B, int N,
                              int d, int relu, int bn, gp_stream_t stream) {
  GP_REQUIRE(y && h && B > 0 && N > 0 && d > 0 && ldh >= d, "relu_bn_fwd: bad args");
  GP_REQUIRE(!bn || (mean && invstd), "relu_bn_fwd: bn needs mean/invstd");
  const int total = B * d;
  const size_t cache_bytes = (size_t)total * sizeof(float);
  if (bn && cache_bytes > 16 * 1024 && cache_bytes <= kNodeCacheMax) {
    auto kern = relu_bn_fwd_kernel<true>;
    GP_CONFIG_ONCE(GP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNodeCacheMax)));
    kern<<<N, 1024, cache_bytes, S(stream)>>>(y, h, ldh, mean, invstd, B, N, d, relu, bn);
  } else {
    const int threads = total >= 4096 ? 512 : (total >= 512 ? 256 : 128);
    relu_bn_fwd_kernel<false><<<N, threads, 0, S(stream)>>>(y, h, ldh, mean, invstd, B, N, d, relu, bn);
  }
  GP_LAUNCHED();
  return GP_OK;
}

namespace gp {
// Generic (any d, any alignment) path of gp_gcn_layer_bwd_x; the vectorised fast paths live in layer_bwd.cu.
int layer_bwd_generic(const gp_layer_bwd* q, cudaStream_t st) {
  const int B = q->B, N = q->N, d = q->d;
  // ws layout of this path: [N * 128 floats: batch-split partials][B*N*d: borrowed fp32 dV (only when the caller
  // wants db without dv)][column-sum scratch]
  float* split_ws = q->ws;
  float* rest = q->ws != nullptr ? q->ws + (long long)N * 128 : nullptr;
  float* dv = q->dv;
  float* cs_ws = rest;
  if (dv == nullptr && q->db != nullptr) {               // column sums need an fp32 dV: borrow it from ws
    dv = rest;
    cs_ws = rest + (long long)B * N * d;
  }
  LbG a;
  a.dz = q->dz; a.lddz = q->lddz; a.dxn = q->dxn; a.lddxn = q->lddxn > 0 ? q->lddxn : (long long)d;
  a.dout = q->dout; a.argidx = q->argidx; a.ldo = q->ldo;
  a.h = q->h; a.ldh = q->ldh; a.y = q->y; a.ldy = q->ldy;
  a.rnorm = q->rnorm; a.mean = q->mean; a.invstd = q->invstd;
  a.B = B; a.N = N; a.d = d; a.relu = q->relu; a.bn = q->bn; a.normalize = q->normalize;
  a.dv = dv; a.dvb = reinterpret_cast<__nv_bfloat16*>(q->dv_bf16); a.lddvb = q->lddvb;
  const int total = B * d;
  const int maxe = d <= 128 ? 4 : (d <= 512 ? 16 : 0);
  const int S = split_ws != nullptr ? split_slices(B, N, d) : 0;
  if (S > 0) {
    const int bs = (B + S - 1) / S;
    dim3 grid(N, S);
    if (q->bn) {
      layer_bwd_split_stats_kernel<<<grid, 256, 0, st>>>(a, bs, split_ws);
      GP_LAUNCHED();
    }
    if (maxe == 4) layer_bwd_split_apply_kernel<4><<<grid, 256, 0, st>>>(a, bs, split_ws);
    else if (maxe == 16) layer_bwd_split_apply_kernel<16><<<grid, 256, 0, st>>>(a, bs, split_ws);
    else layer_bwd_split_apply_kernel<0><<<grid, 256, 0, st>>>(a, bs, split_ws);
    GP_LAUNCHED();
  } else {
    const size_t cache_bytes = (size_t)total * sizeof(float);
    const bool use_cache = q->bn && cache_bytes > 16 * 1024 && cache_bytes <= kNodeCacheMax;
    // 512 threads: the MAXE=16 variant needs 80 registers/thread (1024 threads would exceed the register file)
    const int threads = use_cache ? 512 : (total >= 4096 ? 512 : (total >= 512 ? 256 : 128));
#define GP_LAUNCH_LBWD(C_, E_)                                                                              \
  do {                                                                                                     \
    auto kern = gcn_layer_bwd_kernel<C_, E_>;                                                              \
    if (C_) {                                                                                              \
      GP_CONFIG_ONCE(GP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNodeCacheMax))); \
    }                                                                                                      \
    kern<<<N, threads, C_ ? cache_bytes : 0, st>>>(a);                                                     \
  } while (0)
    if (use_cache) {
      if (maxe == 4) GP_LAUNCH_LBWD(true, 4); else if (maxe == 16) GP_LAUNCH_LBWD(true, 16); else GP_LAUNCH_LBWD(true, 0);
    } else {
      if (maxe == 4) GP_LAUNCH_LBWD(false, 4); else if (maxe == 16) GP_LAUNCH_LBWD(false, 16); else GP_LAUNCH_LBWD(false, 0);
    }
#undef GP_LAUNCH_LBWD
    GP_LAUNCHED();
  }
  if (q->db != nullptr) GP_TRY(colsum(dv, (long long)B * N, d, d, q->db, 0, cs_ws, st));
  return GP_OK;
}
}  // namespace gp

extern "C" int gp_gcn_layer_bwd(const float* dz, long long lddz, const float* dxn, const float* dout,
                                const int32_t* argidx, long long ldo, const float* h, long long ldh,
                                const float* y, long long ldy, const float* rnorm, const float* invstd,
                                int B, int N, int d, int relu, int bn, int normalize, float* dv,
                                gp_stream_t stream) {
  gp_layer_bwd q;
  q.dz = dz; q.lddz = lddz; q.dxn = dxn; q.dout = dout; q.argidx = argidx; q.ldo = ldo;
  q.h = h; q.ldh = ldh; q.y = y; q.ldy = ldy; q.rnorm = rnorm; q.mean = nullptr; q.invstd = invstd;
  q.B = B; q.N = N; q.d = d; q.relu = relu; q.bn = bn; q.normalize = normalize;
  q.dv = dv; q.dv_bf16 = nullptr; q.lddvb = 0; q.db = nullptr; q.ws = nullptr; q.lddxn = 0; q.nb_zero = nullptr;
  q.dz_bf16 = q.dxn_bf16 = 0;
  return gp_gcn_layer_bwd_x(&q, stream);
}

extern "C" int gp_readout_max_fwd(const float* z, long long ldz, const int32_t* nb, int B, int N, int F,
                                  float* out, int32_t* argidx, long long ldo, gp_stream_t stream) {
  GP_REQUIRE(z && out && argidx && B > 0 && N > 0 && F > 0, "readout_max_fwd: bad args");
  GP_REQUIRE(B <= 65535, "readout_max_fwd: B too large");
  dim3 grid((F + 31) / 32, B), block(32, 8);
  readout_max_kernel<<<grid, block, 0, S(stream)>>>(z, ldz, nb, N, N, F, out, argidx, ldo);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_readout_max_fwd_x(const float* z, long long ldz, int pitch, const int32_t* nb, int B, int N, int F,
                                    float* out, int32_t* argidx, long long ldo, gp_stream_t stream) {
  GP_REQUIRE(z && out && argidx && B > 0 && N > 0 && F > 0 && pitch >= N, "readout_max_fwd_x: bad args");
  GP_REQUIRE(B <= 65535, "readout_max_fwd_x: B too large");
  dim3 grid((F + 31) / 32, B), block(32, 8);
  readout_max_kernel<<<grid, block, 0, S(stream)>>>(z, ldz, nb, N, pitch, F, out, argidx, ldo);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_softmax_mask_fwd(float* t, const int32_t* nb, int B, int N, int K, gp_stream_t stream) {
  GP_REQUIRE(t && B > 0 && N > 0 && K > 0, "softmax_mask_fwd: bad args");
  const long long rows = (long long)B * N;
  softmax_mask_fwd_kernel<<<warp_grid(rows, 256), 256, 0, S(stream)>>>(t, nb, rows, N, K);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_softmax_mask_bwd(const float* s, const float* ds, const int32_t* nb, int B, int N, int K,
                                   float* dt, gp_stream_t stream) {
  GP_REQUIRE(s && ds && dt && B > 0 && N > 0 && K > 0, "softmax_mask_bwd: bad args");
  const long long rows = (long long)B * N;
  softmax_mask_bwd_kernel<<<warp_grid(rows, 256), 256, 0, S(stream)>>>(s, ds, nb, rows, N, K, dt);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_ce_fwd(const float* logits, const int64_t* label, int B, int C, float* loss, float* probs,
                         gp_stream_t stream) {
  GP_REQUIRE(logits && label && loss && probs && B > 0 && C > 0, "ce_fwd: bad args");
  ce_fwd_kernel<<<1, 256, 0, S(stream)>>>(logits, label, B, C, loss, probs);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_ce_bwd(const float* probs, const int64_t* label, const float* upstream, int B, int C,
                         float* dlogits, gp_stream_t stream) {
  GP_REQUIRE(probs && label && dlogits && B > 0 && C > 0, "ce_bwd: bad args");
  const int total = B * C;
  ce_bwd_kernel<<<(total + 255) / 256, 256, 0, S(stream)>>>(probs, label, upstream, B, C, dlogits);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_colsum_f32(const float* x, long long rows, int d, long long ld, float* out, int accumulate,
                             float* ws, gp_stream_t stream) {
  return colsum(x, rows, d, ld, out, accumulate, ws, S(stream));
}

extern "C" int gp_relu_mask_bwd(const float* dy, const float* y, long long n, float* dx, gp_stream_t stream) {
  GP_REQUIRE(dy && y && dx && n > 0, "relu_mask_bwd: bad args");
  long long blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  relu_mask_bwd_kernel<<<(int)blocks, 256, 0, S(stream)>>>(dy, y, n, dx);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_bias_normalize_f32(float* v, const float* bias, float* rnorm, long long rows, int d,
                                     long long ld, int normalize, gp_stream_t stream) {
  GP_REQUIRE(v && rows > 0 && d > 0 && ld >= d, "bias_normalize: bad args");
  GP_REQUIRE(!normalize || rnorm, "bias_normalize: normalize needs rnorm");
  return bias_normalize(v, bias, rnorm, rows, d, ld, normalize, S(stream));
}

extern "C" int gp_fill_f32(float* x, long long n, float v, gp_stream_t stream) {
  GP_REQUIRE(x && n >= 0, "fill: bad args");
  if (n == 0) return GP_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  fill_kernel<<<(int)blocks, 256, 0, S(stream)>>>(x, n, v);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_fill_i32(int32_t* x, long long n, int32_t v, gp_stream_t stream) {
  GP_REQUIRE(x && n > 0, "fill_i32: bad args");
  long long blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  fill_i32_kernel<<<(int)blocks, 256, 0, S(stream)>>>(x, n, v);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_pad_copy_f32(const float* src, long long ld_src, long long rows, int cols, float* dst,
                               long long ld_dst, long long rows_dst, int cols_dst, float fill, gp_stream_t stream) {
  GP_REQUIRE(src && dst && rows > 0 && cols > 0 && rows_dst >= rows && cols_dst >= cols && ld_src >= cols &&
             ld_dst >= cols_dst, "pad_copy: bad args");
  long long blocks = (rows_dst * cols_dst + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  pad_copy_kernel<<<(int)blocks, 256, 0, S(stream)>>>(src, ld_src, rows, cols, dst, ld_dst, rows_dst, cols_dst, fill);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_dropout_f32(const float* x, long long ldx, long long rows, int d, float p, unsigned long long seed,
                              float* y, long long ldy, void* y_bf16, long long ldyb, gp_stream_t stream) {
  GP_REQUIRE(x && (y || y_bf16) && rows > 0 && d > 0 && ldx >= d && p >= 0.f && p < 1.f, "dropout: bad args");
  GP_REQUIRE((!y || ldy >= d) && (!y_bf16 || ldyb >= d), "dropout: bad output strides");
  long long blocks = (rows * d + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  dropout_kernel<<<(int)blocks, 256, 0, S(stream)>>>(x, ldx, rows, d, p, seed, y, ldy,
                                                    reinterpret_cast<__nv_bfloat16*>(y_bf16), ldyb);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_axpy_f32(const float* x, float* y, long long n, float a, gp_stream_t stream) {
  GP_REQUIRE(x && y && n > 0, "axpy: bad args");
  long long blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  axpy_kernel<<<(int)blocks, 256, 0, S(stream)>>>(x, y, n, a);
  GP_LAUNCHED();
  return GP_OK;
}
