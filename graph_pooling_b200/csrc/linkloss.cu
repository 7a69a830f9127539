// Fused link-prediction loss (encoders.py:1311-1331): 64x64 tiles of P = S S^T are formed in
// registers, compared with the adjacency tile (masked BCE), block-reduced into one partial per
// tile, and -- for training -- the symmetrised gradient dl/dP + (dl/dP)^T is written once so that
// the backward is a single GEMM gsym.S.  P itself never touches HBM (the reference materialises
// ~6 [B,N,N] temporaries here).  Tiles beyond a graph's node count exit immediately.
#include <cuda_bf16.h>
#include "common.cuh"

namespace gp {

constexpr float kEpsLink = 1e-7f;

__device__ __forceinline__ void bce(float a, float p, bool over, float& l, float& g) {
  l = -a * logf(p + kEpsLink) - (1.f - a) * logf(1.f - p + kEpsLink);
  g = over ? 0.f : (-a / (p + kEpsLink) + (1.f - a) / (1.f - p + kEpsLink));
}

__global__ void __launch_bounds__(256)
linkloss_fwd_kernel(const float* __restrict__ s, const float* __restrict__ adj, const int32_t* __restrict__ nb,
                    int N, int K, float* __restrict__ partial, float* __restrict__ gsym) {
  constexpr int BT = 64, BK = 16;
  __shared__ __align__(16) float Si[BK][BT + 4];
  __shared__ __align__(16) float Sj[BK][BT + 4];
  __shared__ float At[BT][BT + 1];
  __shared__ float sh[33];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int b = blockIdx.z, i0 = blockIdx.y * BT, j0 = blockIdx.x * BT;
  const int T = gridDim.x;
  const int nreal = nb != nullptr ? min(nb[b], N) : N;
  const long long pidx = ((long long)b * T + blockIdx.y) * T + blockIdx.x;
  if (i0 >= nreal || j0 >= nreal) {
    if (tid == 0) partial[pidx] = 0.f;
    return;
  }
  const float* sb = s + (long long)b * N * K;
  const float* ab = adj + (long long)b * N * N;

  // P = <S_i, S_j> is summed per 16-term tile and the tile partials are added with Kahan's correction: the error of
  // P (and of G = dl/dP, which divides by P) stays at ~1e-7 for any cluster count (see bgemm_simt.cu, COMP)
  float acc[4][4], comp[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; comp[i][j] = 0.f; }

  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      const int k = e % BK, m = e / BK;
      const int gk = k0 + k;
      Si[k][m] = (i0 + m < nreal && gk < K) ? sb[(long long)(i0 + m) * K + gk] : 0.f;
      Sj[k][m] = (j0 + m < nreal && gk < K) ? sb[(long long)(j0 + m) * K + gk] : 0.f;
    }
    __syncthreads();
    float part[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], c[4];
      a[0] = Si[kk][ty * 2]; a[1] = Si[kk][ty * 2 + 1]; a[2] = Si[kk][32 + ty * 2]; a[3] = Si[kk][32 + ty * 2 + 1];
      c[0] = Sj[kk][tx * 2]; c[1] = Sj[kk][tx * 2 + 1]; c[2] = Sj[kk][32 + tx * 2]; c[3] = Sj[kk][32 + tx * 2 + 1];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a[i], c[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float y = part[i][j] - comp[i][j];
        const float t = acc[i][j] + y;
        comp[i][j] = (t - acc[i][j]) - y;
        acc[i][j] = t;
      }
    __syncthreads();
  }

  // transposed adjacency tile A[j0+n, i0+m] staged through shared memory (coalesced read)
  if (gsym != nullptr) {
    for (int e = tid; e < BT * BT; e += 256) {
      const int r = e / BT, c = e % BT;            // r: row in block j, c: column in block i
      At[r][c] = (j0 + r < nreal && i0 + c < nreal) ? ab[(long long)(j0 + r) * N + (i0 + c)] : 0.f;
    }
    __syncthreads();
  }

  float lsum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int lm = (i < 2 ? ty * 2 + i : 32 + ty * 2 + (i - 2));
    const int m = i0 + lm;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ln = (j < 2 ? tx * 2 + j : 32 + tx * 2 + (j - 2));
      const int n = j0 + ln;
      if (m < nreal && n < nreal) {
        float p = acc[i][j];
        const bool over = p > 1.f;                 // R3: clamp(max=1), zero gradient beyond it
        if (over) p = 1.f;
        const float a = ab[(long long)m * N + n];
        float l, g;
        bce(a, p, over, l, g);
        lsum += l;
        if (gsym != nullptr) {
          float l2, g2;
          bce(At[ln][lm], p, over, l2, g2);
          gsym[((long long)b * N + m) * N + n] = g + g2;
        }
      }
    }
  }
  const float tot = block_sum(lsum, sh);
  if (tid == 0) partial[pidx] = tot;
}

// stage 1 of the finalisation for large partial arrays: R blocks -> R sums appended after the array
__global__ void loss_partial_reduce_kernel(float* __restrict__ partial, int n_partial, int R) {
  __shared__ double shd[256];
  const int per = (n_partial + R - 1) / R;
  const int i0 = blockIdx.x * per, i1 = min(n_partial, i0 + per);
  double s = 0.0;
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) s += (double)partial[i];
  shd[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) shd[threadIdx.x] += shd[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[n_partial + blockIdx.x] = (float)shd[0];
}

__global__ void loss_finalize_kernel(const float* __restrict__ partial, int n_partial, double inv_entries,
                                     const float* __restrict__ ce, float* __restrict__ total,
                                     float* __restrict__ link) {
  __shared__ double shd[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n_partial; i += blockDim.x) s += (double)partial[i];
  shd[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) shd[threadIdx.x] += shd[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float l = (float)(shd[0] * inv_entries);
    if (link != nullptr) *link = l;
    if (total != nullptr) *total = (ce != nullptr ? *ce : 0.f) + l;
  }
}

// ---------------------------------------------------------------------------------------------
// adj_hop > 1 (encoders.py:1312-1317): the predicted adjacency is Q = sum_{h=1..hop} (S S^T)^h, materialised by the
// caller's GEMM chain; this pass does the clamp (R3: min(Q, 1)), the masked BCE reduction and writes G = dl/dQ
// (fp32, NOT symmetrised: the backward pushes it through the powers of P first), zero where the clamp is active
// or outside the n_b x n_b block.  One partial per 32-row x 256-column strip.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
linkloss_from_q_kernel(const float* __restrict__ Q, const float* __restrict__ adj, const int32_t* __restrict__ nb,
                       int N, float* __restrict__ partial, float* __restrict__ G) {
  __shared__ float sh[33];
  const int b = blockIdx.z, i0 = blockIdx.y * 32, j0 = blockIdx.x * 256;
  const int nreal = nb != nullptr ? min(nb[b], N) : N;
  const long long pidx = ((long long)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const float* ab = adj + (long long)b * N * N;
  const float* qb = Q + (long long)b * N * N;
  float lsum = 0.f;
  const int n = j0 + threadIdx.x;
  if (n < N) {
    for (int r = 0; r < 32; ++r) {
      const int m = i0 + r;
      if (m >= N) break;
      float g = 0.f;
      if (m < nreal && n < nreal) {
        float q = qb[(long long)m * N + n];
        const bool over = q > 1.f;
        if (over) q = 1.f;
        float l;
        bce(ab[(long long)m * N + n], q, over, l, g);
        lsum += l;
      }
      if (G != nullptr) G[((long long)b * N + m) * N + n] = g;
    }
  }
  const float tot = block_sum(lsum, sh);
  if (threadIdx.x == 0) partial[pidx] = tot;
}

}  // namespace gp

using namespace gp;

extern "C" int gp_linkloss_from_q_partials(int B, int N) { return B * ((N + 31) / 32) * ((N + 255) / 256); }
extern "C" int gp_linkloss_from_q(const float* Q, const float* adj, const int32_t* nb, int B, int N, float* partial,
                                  float* G, gp_stream_t stream) {
  GP_REQUIRE(Q && adj && partial && B > 0 && N > 0, "linkloss_from_q: bad args");
  const int Ty = (N + 31) / 32, Tx = (N + 255) / 256;
  GP_REQUIRE(B <= 65535 && Ty <= 65535, "linkloss_from_q: grid too large");
  dim3 grid(Tx, Ty, B);
  linkloss_from_q_kernel<<<grid, 256, 0, S(stream)>>>(Q, adj, nb, N, partial, G);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_linkloss_fwd(const float* s, const float* adj, const int32_t* nb, int B, int N, int K,
                               float* partial, float* gsym, gp_stream_t stream) {
  GP_REQUIRE(s && adj && partial && B > 0 && N > 0 && K > 0, "linkloss_fwd: bad args");
  const int T = (N + 63) / 64;
  GP_REQUIRE(B <= 65535 && T <= 65535, "linkloss_fwd: grid too large");
  dim3 grid(T, T, B);
  linkloss_fwd_kernel<<<grid, 256, 0, S(stream)>>>(s, adj, nb, N, K, partial, gsym);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_loss_finalize(const float* partial, int n_partial, double inv_entries, const float* ce,
                                float* total, float* link, gp_stream_t stream) {
  GP_REQUIRE(partial && n_partial > 0, "loss_finalize: bad args");
  if (n_partial > 8192) {     // two stages; the caller provides 256 floats of scratch after the array
    loss_partial_reduce_kernel<<<256, 256, 0, S(stream)>>>(const_cast<float*>(partial), n_partial, 256);
    GP_LAUNCHED();
    partial += n_partial;
    n_partial = 256;
  }
  loss_finalize_kernel<<<1, 256, 0, S(stream)>>>(partial, n_partial, inv_entries, ce, total, link);
  GP_LAUNCHED();
  return GP_OK;
}
