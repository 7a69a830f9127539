// North-star loss options that the reference does NOT contain (BASELINE.json north_star):
//   Frobenius link loss   L_F = mean_b || (A_b - S_b S_b^T) restricted to the n_b x n_b block ||_F
//   row entropy           L_E = (1 / sum_b n_b) * sum_{b, n < n_b} - sum_k S log(S + eps)
// Oracle: oracle/diffpool_oracle.py frobenius_link_loss / row_entropy_loss (the DiffPool paper's definitions;
// parity unpinned by the reference).  fp32 kernels here; the tensor-core Frobenius epilogue is in gemm_tc2.cu.
#include <cuda_bf16.h>
#include "common.cuh"

namespace gp {

constexpr float kEpsEnt = 1e-7f;

// 64x64 tiles of P = S S^T in registers (same tiling as linkloss.cu); partial[b][ty][tx] = sum d^2 over the tile,
// gsym = -((A - P) + (A^T - P)) = dL/dP + (dL/dP)^T up to the per-graph factor 1 / (B * ||d_b||_F).
__global__ void __launch_bounds__(256)
frob_fwd_kernel(const float* __restrict__ s, const float* __restrict__ adj, const int32_t* __restrict__ nb,
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
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
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
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], c[4];
      a[0] = Si[kk][ty * 2]; a[1] = Si[kk][ty * 2 + 1]; a[2] = Si[kk][32 + ty * 2]; a[3] = Si[kk][32 + ty * 2 + 1];
      c[0] = Sj[kk][tx * 2]; c[1] = Sj[kk][tx * 2 + 1]; c[2] = Sj[kk][32 + tx * 2]; c[3] = Sj[kk][32 + tx * 2 + 1];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], c[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (gsym != nullptr) {
    for (int e = tid; e < BT * BT; e += 256) {
      const int r = e / BT, c = e % BT;
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
        const float p = acc[i][j];
        const float d = ab[(long long)m * N + n] - p;
        lsum = fmaf(d, d, lsum);
        if (gsym != nullptr) gsym[((long long)b * N + m) * N + n] = -(d + (At[ln][lm] - p));
      }
    }
  }
  const float tot = block_sum(lsum, sh);
  if (tid == 0) partial[pidx] = tot;
}

// one block per graph: norm[b] = sqrt(sum of the graph's partials), coef[b] = 1 / (B * norm[b]) (0 if norm == 0)
__global__ void frob_graph_norm_kernel(const float* __restrict__ partial, int per_graph, int B,
                                       float* __restrict__ norm, float* __restrict__ coef) {
  __shared__ double shd[256];
  const int b = blockIdx.x;
  const float* p = partial + (long long)b * per_graph;
  double s = 0.0;
  for (int i = threadIdx.x; i < per_graph; i += blockDim.x) s += (double)p[i];
  shd[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) shd[threadIdx.x] += shd[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float nrm = (float)sqrt(shd[0]);
    norm[b] = nrm;
    coef[b] = nrm > 0.f ? 1.f / ((float)B * nrm) : 0.f;
  }
}

// out[r, :] = scale[b(r)] * (*upstream) * x[r, :]   (fp32 and/or bf16 copy, row strides ldo / ldob)
__global__ void scale_rows_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                  const float* __restrict__ upstream, long long rows, int rows_per_batch, int cols,
                                  float* __restrict__ out, long long ldo, __nv_bfloat16* __restrict__ outb,
                                  long long ldob, int cols_pad) {
  const float up = upstream != nullptr ? *upstream : 1.f;
  const long long total = rows * cols_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols_pad;
    const int c = (int)(i - r * cols_pad);
    float v = 0.f;
    if (c < cols) v = x[r * cols + c] * scale[r / rows_per_batch] * up;
    if (out != nullptr && c < cols) out[r * ldo + c] = v;
    if (outb != nullptr) outb[r * ldob + c] = __float2bfloat16_rn(v);
  }
}

// one warp per row: e = -sum_k s log(s + eps) on real rows; per-block partial sums (deterministic)
__global__ void __launch_bounds__(256)
entropy_fwd_kernel(const float* __restrict__ s, const int32_t* __restrict__ nb, long long rows, int N, int K,
                   float* __restrict__ partial) {
  __shared__ float sh[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * 8 + warp, nw = (long long)gridDim.x * 8;
  float acc = 0.f;
  for (long long r = gw; r < rows; r += nw) {
    const int b = (int)(r / N), n = (int)(r - (long long)b * N);
    if (nb != nullptr && n >= nb[b]) continue;
    const float* p = s + r * K;
    float e = 0.f;
    for (int c = lane; c < K; c += 32) {
      const float v = p[c];
      e -= v * logf(v + kEpsEnt);
    }
    acc += e;                                            // lanes hold disjoint column subsets
  }
  acc = warp_sum(acc);
  if (lane == 0) sh[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w];
    partial[blockIdx.x] = t;
  }
}

// ds (+)= upstream * scale * -(log(s + eps) + s / (s + eps)) on real rows (0 on pad rows)
__global__ void entropy_bwd_kernel(const float* __restrict__ s, const int32_t* __restrict__ nb, long long total,
                                   int N, int K, const float* __restrict__ upstream, float scale,
                                   float* __restrict__ ds, int accumulate) {
  const float up = (upstream != nullptr ? *upstream : 1.f) * scale;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / K;
    const int b = (int)(r / N), n = (int)(r - (long long)b * N);
    float g = 0.f;
    if (nb == nullptr || n < nb[b]) {
      const float v = s[i];
      g = -up * (logf(v + kEpsEnt) + v / (v + kEpsEnt));
    }
    ds[i] = accumulate ? ds[i] + g : g;
  }
}

// total = (base ? *base : 0) + w * (*term)
__global__ void add_scaled_kernel(const float* base, const float* term, float w, float* total) {
  *total = (base != nullptr ? *base : 0.f) + w * (*term);
}

// out = (1/B) * sum_b norm[b]  (+ ce)
__global__ void frob_mean_kernel(const float* __restrict__ norm, int B, const float* ce, float* total, float* link) {
  __shared__ double shd[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) s += (double)norm[i];
  shd[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) shd[threadIdx.x] += shd[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float l = (float)(shd[0] / (double)B);
    if (link != nullptr) *link = l;
    if (total != nullptr) *total = (ce != nullptr ? *ce : 0.f) + l;
  }
}

// out[0] = 1 / sum_b n_b^2 ; out[1] = 1 / sum_b n_b   (the link-loss / entropy normalisers when the node counts live
// on the device only, e.g. inside a captured CUDA graph: no host round trip)
__global__ void nb_stats_kernel(const int32_t* __restrict__ nb, int B, float* __restrict__ out) {
  __shared__ double s2[256], s1[256];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const double n = (double)nb[i];
    a += n * n; b += n;
  }
  s2[threadIdx.x] = a; s1[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s2[threadIdx.x] += s2[threadIdx.x + o]; s1[threadIdx.x] += s1[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = s2[0] > 0.0 ? (float)(1.0 / s2[0]) : 0.f;
    out[1] = s1[0] > 0.0 ? (float)(1.0 / s1[0]) : 0.f;
  }
}

// prod = (*a) * (*b) ; total = (c ? *c : 0) + prod
__global__ void mul_add_dev_kernel(const float* a, const float* b, const float* c, float* total, float* prod) {
  const float p = (*a) * (*b);
  if (prod != nullptr) *prod = p;
  if (total != nullptr) *total = (c != nullptr ? *c : 0.f) + p;
}

static int grid_for(long long n, int per_block) {
  long long b = (n + per_block - 1) / per_block;
  if (b > kNumSMs * 8) b = kNumSMs * 8;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace gp

using namespace gp;

extern "C" int gp_frob_link_fwd(const float* s, const float* adj, const int32_t* nb, int B, int N, int K,
                                float* partial, float* gsym, gp_stream_t stream) {
  GP_REQUIRE(s && adj && partial && B > 0 && N > 0 && K > 0, "frob_link_fwd: bad args");
  const int T = (N + 63) / 64;
  GP_REQUIRE(B <= 65535 && T <= 65535, "frob_link_fwd: grid too large");
  dim3 grid(T, T, B);
  frob_fwd_kernel<<<grid, 256, 0, S(stream)>>>(s, adj, nb, N, K, partial, gsym);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_frob_finalize(const float* partial, int per_graph, int B, const float* ce, float* total,
                                float* link, float* norm, float* coef, gp_stream_t stream) {
  GP_REQUIRE(partial && per_graph > 0 && B > 0 && norm && coef, "frob_finalize: bad args");
  frob_graph_norm_kernel<<<B, 256, 0, S(stream)>>>(partial, per_graph, B, norm, coef);
  GP_LAUNCHED();
  frob_mean_kernel<<<1, 256, 0, S(stream)>>>(norm, B, ce, total, link);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_scale_rows_batch(const float* x, const float* scale, const float* upstream, int B,
                                   int rows_per_batch, int cols, float* out, long long ldo, void* out_bf16,
                                   long long ldob, int cols_pad, gp_stream_t stream) {
  GP_REQUIRE(x && scale && (out || out_bf16) && B > 0 && rows_per_batch > 0 && cols > 0, "scale_rows_batch: bad args");
  if (cols_pad < cols) cols_pad = cols;
  GP_REQUIRE(out_bf16 == nullptr || ldob >= cols_pad, "scale_rows_batch: ldob < cols_pad");
  const long long rows = (long long)B * rows_per_batch;
  scale_rows_kernel<<<grid_for(rows * cols_pad, 256), 256, 0, S(stream)>>>(
      x, scale, upstream, rows, rows_per_batch, cols, out, ldo, reinterpret_cast<__nv_bfloat16*>(out_bf16), ldob,
      cols_pad);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_entropy_partials(int B, int N) {
  return grid_for((long long)B * N, 8);
}

extern "C" int gp_entropy_fwd(const float* s, const int32_t* nb, int B, int N, int K, float* partial,
                              gp_stream_t stream) {
  GP_REQUIRE(s && partial && B > 0 && N > 0 && K > 0, "entropy_fwd: bad args");
  const long long rows = (long long)B * N;
  entropy_fwd_kernel<<<grid_for(rows, 8), 256, 0, S(stream)>>>(s, nb, rows, N, K, partial);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_entropy_bwd(const float* s, const int32_t* nb, int B, int N, int K, const float* upstream,
                              float scale, float* ds, int accumulate, gp_stream_t stream) {
  GP_REQUIRE(s && ds && B > 0 && N > 0 && K > 0, "entropy_bwd: bad args");
  const long long total = (long long)B * N * K;
  entropy_bwd_kernel<<<grid_for(total, 256), 256, 0, S(stream)>>>(s, nb, total, N, K, upstream, scale, ds, accumulate);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_add_scaled(const float* base, const float* term, float w, float* total, gp_stream_t stream) {
  GP_REQUIRE(term && total, "add_scaled: bad args");
  add_scaled_kernel<<<1, 1, 0, S(stream)>>>(base, term, w, total);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_nb_stats(const int32_t* nb, int B, float* out, gp_stream_t stream) {
  GP_REQUIRE(nb && out && B > 0, "nb_stats: bad args");
  nb_stats_kernel<<<1, 256, 0, S(stream)>>>(nb, B, out);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_mul_add_dev(const float* a, const float* b, const float* c, float* total, float* prod,
                              gp_stream_t stream) {
  GP_REQUIRE(a && b && (total || prod), "mul_add_dev: bad args");
  mul_add_dev_kernel<<<1, 1, 0, S(stream)>>>(a, b, c, total, prod);
  GP_LAUNCHED();
  return GP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Optimiser step of train.py:209-210 over FLAT buffers: clip_grad_norm(max_norm) folded into Adam (lr, betas, eps;
// no weight decay, no amsgrad -- torch.optim.Adam's defaults, train.py:173).  Deterministic two-stage norm, the step
// counter lives on the device (CUDA-graph friendly), one pass over p / g / m / v.
// ---------------------------------------------------------------------------------------------------------
namespace gp {
__global__ void sumsq_stage1(const float* __restrict__ g, long long n, float* __restrict__ part) {
  __shared__ double shd[256];
  double s = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = g[i];
    s += v * v;
  }
  shd[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) shd[threadIdx.x] += shd[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = (float)shd[0];
}
__global__ void sumsq_stage2(const float* __restrict__ part, int nparts, float* __restrict__ out) {
  __shared__ double shd[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += (double)part[i];
  shd[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) shd[threadIdx.x] += shd[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)shd[0];
}
__global__ void adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                 const float* __restrict__ step_dev, const float* __restrict__ sumsq, float max_norm,
                                 float grad_scale) {
  const float t = *step_dev + 1.f;
  // g' = grad_scale * g (data parallel: 1 / world after a SUM all-reduce), then clip_grad_norm on ||g'||
  float coef = grad_scale;
  if (max_norm > 0.f && sumsq != nullptr) coef *= fminf(1.f, max_norm / (grad_scale * sqrtf(*sumsq) + 1e-6f));
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = lr / bc1, rs2 = rsqrtf(bc2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);             // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * b2 + gi * gi * (1.f - b2);
    m[i] = mi; v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * rs2 + eps);
  }
}
__global__ void bump_kernel(float* step_dev) { *step_dev += 1.f; }
// g *= pre_scale * min(1, max_norm / (pre_scale * sqrt(sumsq) + 1e-6)): clip_grad_norm_ on a flat buffer whose sum of
// squares was taken BEFORE the pre-scale (max_norm <= 0 or sumsq == NULL: pre-scale only)
__global__ void clip_scale_kernel(float* __restrict__ g, long long n, const float* __restrict__ sumsq, float max_norm,
                                  float pre_scale) {
  float coef = pre_scale;
  if (max_norm > 0.f && sumsq != nullptr) coef *= fminf(1.f, max_norm / (pre_scale * sqrtf(*sumsq) + 1e-6f));
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    g[i] *= coef;
}
// dst_e[i] += alpha * src_e[i] for up to 64 (src, dst, n) entries per launch: every parameter gradient of a backward
// pass is added into its slot of the flat gradient buffer by ONE launch (instead of one autograd AccumulateGrad add
// per parameter).  blockIdx.y = entry.
struct AxpyTable { const float* src[64]; float* dst[64]; long long n[64]; };
__global__ void multi_axpy_kernel(const __grid_constant__ AxpyTable t, float alpha) {
  const int e = blockIdx.y;
  const float* __restrict__ s = t.src[e];
  float* __restrict__ d = t.dst[e];
  const long long n = t.n[e];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = fmaf(alpha, s[i], d[i]);
}
}  // namespace gp

extern "C" int gp_sumsq_f32(const float* g, long long n, float* out, float* ws, gp_stream_t stream) {
  GP_REQUIRE(g && out && ws && n > 0, "sumsq: bad args (ws: 1024 floats)");
  const int blocks = grid_for(n, 256 * 8) > 1024 ? 1024 : grid_for(n, 256 * 8);
  gp::sumsq_stage1<<<blocks, 256, 0, S(stream)>>>(g, n, ws);
  GP_LAUNCHED();
  gp::sumsq_stage2<<<1, 256, 0, S(stream)>>>(ws, blocks, out);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_adam_step_f32(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                                float beta2, float eps, float* step_dev, const float* sumsq_dev, float max_norm,
                                float grad_scale, gp_stream_t stream) {
  GP_REQUIRE(p && g && m && v && step_dev && n > 0 && grad_scale > 0.f, "adam_step: bad args");
  gp::adam_flat_kernel<<<grid_for(n, 256 * 4), 256, 0, S(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, step_dev,
                                                                      sumsq_dev, max_norm, grad_scale);
  GP_LAUNCHED();
  gp::bump_kernel<<<1, 1, 0, S(stream)>>>(step_dev);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_clip_scale_f32(float* g, long long n, const float* sumsq_dev, float max_norm, float pre_scale,
                                 gp_stream_t stream) {
  GP_REQUIRE(g && n > 0 && pre_scale > 0.f, "clip_scale: bad args");
  gp::clip_scale_kernel<<<grid_for(n, 256 * 4), 256, 0, S(stream)>>>(g, n, sumsq_dev, max_norm, pre_scale);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_multi_axpy_f32(const gp_axpy_entry* entries, int count, float alpha, gp_stream_t stream) {
  GP_REQUIRE(entries != nullptr && count > 0, "multi_axpy: bad args");
  for (int e0 = 0; e0 < count; e0 += 64) {
    gp::AxpyTable t;
    const int c = count - e0 < 64 ? count - e0 : 64;
    long long nmax = 0;
    for (int i = 0; i < 64; ++i) {
      const gp_axpy_entry& q = entries[e0 + (i < c ? i : 0)];
      GP_REQUIRE(q.src && q.dst && q.n >= 0, "multi_axpy: entry %d: null pointer or negative length", e0 + i);
      t.src[i] = q.src; t.dst[i] = q.dst; t.n[i] = i < c ? q.n : 0;
      if (i < c && q.n > nmax) nmax = q.n;
    }
    if (nmax == 0) continue;
    int bx = (int)((nmax + 256 * 4 - 1) / (256 * 4));
    if (bx > 64) bx = 64;
    gp::multi_axpy_kernel<<<dim3(bx, c), 256, 0, S(stream)>>>(t, alpha);
    GP_LAUNCHED();
  }
  return GP_OK;
}
