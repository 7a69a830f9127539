// Per-graph fused GraphConv for SMALL graphs (ENZYMES-sized: N <= 128, feature widths <= 128), fp32.
//
// At these sizes a whole graph's working set -- the real n_b x n_b adjacency block, X, U, W -- fits in one SM's
// shared memory, and the generic path (three or more launches per layer, each a latency-bound chain of
// global -> shared round trips over a handful of CTAs) is bound by launch and memory latency, not by FLOPs or
// bandwidth.  Here ONE CTA per graph stages its operands once and runs the whole layer:
//   forward   U = A.X ;  V = U.W + b ;  Y = V / max(||V||_2, 1e-12)            (encoders.py:315-328)
//   backward  dU = dV.W^T ;  dX = A^T.dU ;  dA += dU.X^T ;  per-graph partials of dW = U^T.dV and db = colsum(dV)
//             (reduced over the batch by the deterministic column-sum kernels)
// Only the real n_b rows / columns are staged and multiplied (padding-aware); pad rows get their analytic values
// (U = 0, Y = normalize(b)).  The in-shared-memory products use 4 x 4 register tiles.
#include "common.cuh"

namespace gp {

constexpr float kEpsNormS = 1e-12f;
constexpr int kSmallMaxSmem = 200 * 1024;

int colsum(const float* x, long long rows, int d, long long ld, float* out, int accumulate, float* ws, cudaStream_t st);

__host__ __device__ inline int r4i(int v) { return (v + 3) & ~3; }

// Staging: one warp per row, 4-byte cp.async per element -- every load of a CTA is in flight at once and nothing is
// held in registers, so the whole staging phase costs one memory latency instead of a chain of them.
__device__ __forceinline__ void cpa4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// dst[r][c] (row stride ldd) <- src[r][c] (row stride lds) for r < rows, c < cols; zero for r < rows_pad, c < cols_pad
__device__ __forceinline__ void stage_rows(float* dst, int ldd, const float* src, long long lds, int rows, int cols,
                                           int rows_pad, int cols_pad) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = warp; r < rows_pad; r += nw) {
    float* d = dst + r * ldd;
    if (r < rows) {
      const float* sp = src + (long long)r * lds;
      for (int c = lane; c < cols; c += 32) cpa4(d + c, sp + c);
      for (int c = cols + lane; c < cols_pad; c += 32) d[c] = 0.f;
    } else {
      for (int c = lane; c < cols_pad; c += 32) d[c] = 0.f;
    }
  }
}
// dst[c][r] (row stride ldd) <- src[r][c]: transposed staging (reads stay coalesced along c)
__device__ __forceinline__ void stage_rows_t(float* dst, int ldd, const float* src, long long lds, int rows, int cols,
                                             int rows_pad, int cols_pad) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = warp; r < rows_pad; r += nw) {
    if (r < rows) {
      const float* sp = src + (long long)r * lds;
      for (int c = lane; c < cols; c += 32) cpa4(dst + c * ldd + r, sp + c);
      for (int c = cols + lane; c < cols_pad; c += 32) dst[c * ldd + r] = 0.f;
    } else {
      for (int c = lane; c < cols_pad; c += 32) dst[c * ldd + r] = 0.f;
    }
  }
}

// C[M x Nc] (+)= A[M x K] . B[K x Nc], all in shared memory, row-major with leading dimensions lda / ldb / ldc
// (ldb, ldc multiples of 4; A zero-filled up to r4(M) rows, B zero-filled up to r4(Nc) columns).
template <bool ACC = false>
__device__ __forceinline__ void smem_gemm(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
                                          int M, int Nc, int K, float* __restrict__ C, int ldc, const float* bias) {
  const int ntn = (Nc + 3) >> 2, ntm = (M + 3) >> 2;
  for (int t = threadIdx.x; t < ntm * ntn; t += blockDim.x) {
    const int tm = t / ntn, tn = t - tm * ntn;
    const int i0 = tm * 4, c0 = tn * 4;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[r][e] = 0.f;
    const float* a0 = A + i0 * lda;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float4 b4 = *reinterpret_cast<const float4*>(Bm + k * ldb + c0);
      const float av[4] = {a0[k], a0[lda + k], a0[2 * lda + k], a0[3 * lda + k]};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        acc[r][0] = fmaf(av[r], b4.x, acc[r][0]); acc[r][1] = fmaf(av[r], b4.y, acc[r][1]);
        acc[r][2] = fmaf(av[r], b4.z, acc[r][2]); acc[r][3] = fmaf(av[r], b4.w, acc[r][3]);
      }
    }
    float bs[4] = {0.f, 0.f, 0.f, 0.f};
    if (bias != nullptr) {
#pragma unroll
      for (int e = 0; e < 4; ++e) bs[e] = c0 + e < Nc ? bias[c0 + e] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float4 o = make_float4(acc[r][0] + bs[0], acc[r][1] + bs[1], acc[r][2] + bs[2], acc[r][3] + bs[3]);
      float4* cp = reinterpret_cast<float4*>(C + (i0 + r) * ldc + c0);
      if (ACC) { const float4 old = *cp; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
      *cp = o;
    }
  }
}

// Same product written straight to GLOBAL memory: C[i][c] (+)= ... for i < M, c < Nc (row stride ldc, any alignment)
template <bool ACC>
__device__ __forceinline__ void smem_gemm_to_global(const float* __restrict__ A, int lda, const float* __restrict__ Bm,
                                                    int ldb, int M, int Nc, int K, float* __restrict__ C, long long ldc) {
  const int ntn = (Nc + 3) >> 2, ntm = (M + 3) >> 2;
  for (int t = threadIdx.x; t < ntm * ntn; t += blockDim.x) {
    const int tm = t / ntn, tn = t - tm * ntn;
    const int i0 = tm * 4, c0 = tn * 4;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[r][e] = 0.f;
    const float* a0 = A + i0 * lda;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float4 b4 = *reinterpret_cast<const float4*>(Bm + k * ldb + c0);
      const float av[4] = {a0[k], a0[lda + k], a0[2 * lda + k], a0[3 * lda + k]};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        acc[r][0] = fmaf(av[r], b4.x, acc[r][0]); acc[r][1] = fmaf(av[r], b4.y, acc[r][1]);
        acc[r][2] = fmaf(av[r], b4.z, acc[r][2]); acc[r][3] = fmaf(av[r], b4.w, acc[r][3]);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (i0 + r < M && c0 + e < Nc) {
          float* o = C + (long long)(i0 + r) * ldc + c0 + e;
          *o = ACC ? *o + acc[r][e] : acc[r][e];
        }
  }
}

struct SmallFwd {
  const float* x; long long ldx; const float* adj; const float* w; const float* bias; const int32_t* nb;
  int B, N, din, dout, normalize;
  float* u; float* y; long long ldy; float* rnorm;
};

// shared-memory floats of the forward kernel for max node count N
__host__ __device__ inline size_t small_fwd_floats(int N, int din, int dout) {
  const int n4 = r4i(N), dp = r4i(din), op = r4i(dout);
  const int wide = dp > op ? dp : op;
  return (size_t)n4 * (N + 1) + (size_t)n4 * wide + (size_t)n4 * dp + (size_t)din * op + op;
}

__global__ void __launch_bounds__(256) gconv_small_fwd_kernel(const SmallFwd p) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N, din = p.din, dout = p.dout;
  const int n = p.nb != nullptr ? min(max(p.nb[b], 0), N) : N;
  const int n4 = r4i(n), dp = r4i(din), op = r4i(dout), lda = N + 1;
  const int wide = dp > op ? dp : op;
  float* As = sm;                                        // [r4(N)][N+1]   real block, rows zero-filled to n4
  float* XV = As + (size_t)r4i(N) * lda;                 // [r4(N)][wide]  X (phase 1), then V (phase 2)
  float* Us = XV + (size_t)r4i(N) * wide;                // [r4(N)][dp]
  float* Ws = Us + (size_t)r4i(N) * dp;                  // [din][op]
  float* bs = Ws + (size_t)din * op;                     // [op]
  const float* ab = p.adj + (long long)b * N * N;
  const float* xb = p.x + (long long)b * N * p.ldx;
  // ---- stage A (real block), X (real rows), W, bias ----
  stage_rows(As, lda, ab, N, n, n, n4, n);
  stage_rows(XV, wide, xb, p.ldx, n, din, n, dp);
  stage_rows(Ws, op, p.w, dout, din, dout, din, op);
  for (int c = tid; c < op; c += 256) bs[c] = (p.bias != nullptr && c < dout) ? p.bias[c] : 0.f;
  cpa_wait_all();
  __syncthreads();
  // ---- U = A.X (rows < n), kept in shared memory and written out; pad rows of u are zero ----
  smem_gemm(As, lda, XV, wide, n, din, n, Us, dp, nullptr);
  __syncthreads();
  float* ub = p.u + (long long)b * N * din;
  for (int e = tid; e < N * din; e += 256) {
    const int i = e / din, c = e - i * din;
    ub[e] = i < n ? Us[i * dp + c] : 0.f;
  }
  // ---- V = U.W + b (rows < n) into XV ----
  smem_gemm(Us, dp, Ws, op, n, dout, din, XV, wide, bs);
  __syncthreads();
  // ---- Y = V / max(||V||, eps); pad rows carry V = b; one warp per row ----
  float* yb = p.y + (long long)b * N * p.ldy;
  for (int i = warp; i < N; i += 8) {
    const float* v = i < n ? XV + i * wide : bs;
    float ss = 0.f;
    for (int c = lane; c < dout; c += 32) ss = fmaf(v[c], v[c], ss);
    ss = warp_sum(ss);
    float r = 1.f;
    if (p.normalize) {
      r = fmaxf(sqrtf(ss), kEpsNormS);
      if (lane == 0) p.rnorm[(long long)b * N + i] = r;
    }
    const float inv = 1.f / r;
    for (int c = lane; c < dout; c += 32) yb[(long long)i * p.ldy + c] = p.normalize ? v[c] * inv : v[c];
  }
}

struct SmallBwd {
  const float* dv; const float* u; const float* x; long long ldx; const float* adj; const float* w;
  const int32_t* nb; int B, N, din, dout;
  float* part;            // [B][din*dout + dout]: per-graph dW and db partials
  float* dx;              // [B,N,din] or NULL
  float* dadj;            // [B,N,N] (+=) or NULL
};

__host__ __device__ inline size_t small_bwd_floats(int N, int din, int dout, int need_dx, int need_da) {
  const int n4 = r4i(N), dp = r4i(din), op = r4i(dout);
  size_t f = (size_t)n4 * op            // dVs   [n4][op]
             + (size_t)r4i(din) * r4i(N) // Ut    [dp][r4(N)]   (U transposed: dW = U^T dV)
             + op;                       // column sums
  if (need_dx || need_da) f += (size_t)dout * dp + (size_t)n4 * dp;          // Wt [dout][dp], dUs [n4][dp]
  if (need_dx) f += (size_t)n4 * (N + 4);                                    // At [n4][r4(N)+..] (A transposed)
  if (need_da) f += (size_t)din * r4i(N);                                    // Xt [din][r4(N)]
  return f;
}

__global__ void __launch_bounds__(256) gconv_small_bwd_kernel(const SmallBwd p, int need_dx, int need_da) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int N = p.N, din = p.din, dout = p.dout;
  const int n = p.nb != nullptr ? min(max(p.nb[b], 0), N) : N;
  const int n4 = r4i(n), N4 = r4i(N), dp = r4i(din), op = r4i(dout);
  float* dVs = sm;                                       // [r4(N)][op]   rows < n (zero-filled to n4)
  float* Ut = dVs + (size_t)N4 * op;                     // [dp][N4]      Ut[c][i] = U[i][c]
  float* cs = Ut + (size_t)dp * N4;                      // [op]
  float* Wt = cs + op;                                   // [dout][dp]    Wt[o][c] = W[c][o]
  float* dUs = Wt + ((need_dx || need_da) ? (size_t)dout * dp : 0);           // [N4][dp]
  float* At = dUs + ((need_dx || need_da) ? (size_t)N4 * dp : 0);             // [N4][N+4]  At[j][i] = A[i][j]
  float* Xt = At + (need_dx ? (size_t)N4 * (N + 4) : 0);                      // [din][N4]  Xt[c][j] = X[j][c]
  const int ldat = N + 4;
  const float* dvb = p.dv + (long long)b * N * dout;
  const float* ub = p.u + (long long)b * N * din;
  // ---- db partial: column sums of dV over ALL N rows (pad rows carry gradient through the BatchNorm) ----
  stage_rows(dVs, op, dvb, dout, n, dout, n4, op);
  stage_rows_t(Ut, N4, ub, din, n, din, n, dp);            // Ut[c][i] = U[i][c] (i < n), zero rows for c >= din
  if (need_dx || need_da) stage_rows_t(Wt, dp, p.w, dout, din, dout, dp, dout);   // Wt[o][c] = W[c][o]
  if (need_dx) stage_rows_t(At, ldat, p.adj + (long long)b * N * N, N, n, n, n, n4);   // At[j][i] = A[i][j]
  if (need_da) stage_rows_t(Xt, N4, p.x + (long long)b * N * p.ldx, p.ldx, n, din, N4, din);   // Xt[c][j] = X[j][c]
  cpa_wait_all();
  __syncthreads();
  float* part = p.part + (long long)b * ((long long)din * dout + dout);
  {   // db: one thread per column over all N rows (coalesced across the warp)
    for (int c = tid; c < dout; c += 256) {
      float s = 0.f;
      for (int i = 0; i < N; ++i) s += dvb[(long long)i * dout + c];
      part[(long long)din * dout + c] = s;
    }
  }
  // ---- dW partial = U^T dV : [din x dout] = Ut [dp x n] . dVs [n x op]; tiles written straight to global ----
  {
    const int ntn = op >> 2, ntm = dp >> 2;
    for (int t = tid; t < ntm * ntn; t += 256) {
      const int tm = t / ntn, tn = t - tm * ntn, i0 = tm * 4, c0 = tn * 4;
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[r][e] = 0.f;
      for (int k = 0; k < n; ++k) {
        const float4 b4 = *reinterpret_cast<const float4*>(dVs + k * op + c0);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float a = Ut[(i0 + r) * N4 + k];
          acc[r][0] = fmaf(a, b4.x, acc[r][0]); acc[r][1] = fmaf(a, b4.y, acc[r][1]);
          acc[r][2] = fmaf(a, b4.z, acc[r][2]); acc[r][3] = fmaf(a, b4.w, acc[r][3]);
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (i0 + r < din && c0 + e < dout) part[(long long)(i0 + r) * dout + c0 + e] = acc[r][e];
    }
  }
  if (!need_dx && !need_da) return;
  // ---- dU = dV.W^T (rows < n) ----
  smem_gemm(dVs, op, Wt, dp, n, din, dout, dUs, dp, nullptr);
  __syncthreads();
  if (need_dx) {          // dX = A^T.dU : rows j < n ; pad rows of dx are zero (A's pad columns are zero)
    float* dxb = p.dx + (long long)b * N * din;
    const int ntn = dp >> 2, ntm = n4 >> 2;
    for (int t = tid; t < ntm * ntn; t += 256) {
      const int tm = t / ntn, tn = t - tm * ntn, j0 = tm * 4, c0 = tn * 4;
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[r][e] = 0.f;
      for (int i = 0; i < n; ++i) {
        const float4 b4 = *reinterpret_cast<const float4*>(dUs + i * dp + c0);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float a = At[(j0 + r) * ldat + i];
          acc[r][0] = fmaf(a, b4.x, acc[r][0]); acc[r][1] = fmaf(a, b4.y, acc[r][1]);
          acc[r][2] = fmaf(a, b4.z, acc[r][2]); acc[r][3] = fmaf(a, b4.w, acc[r][3]);
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (j0 + r < n && c0 + e < din) dxb[(long long)(j0 + r) * din + c0 + e] = acc[r][e];
    }
    for (int e = tid + n * din; e < N * din; e += 256) dxb[e] = 0.f;
  }
  if (need_da) {          // dA += dU.X^T : [n x n] += dUs [n x din] . Xt [din x n]
    float* dab = p.dadj + (long long)b * N * N;
    const int ntn = N4 >> 2, ntm = n4 >> 2;
    for (int t = tid; t < ntm * ntn; t += 256) {
      const int tm = t / ntn, tn = t - tm * ntn, i0 = tm * 4, j0 = tn * 4;
      if (j0 >= n) continue;
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[r][e] = 0.f;
      for (int c = 0; c < din; ++c) {
        const float4 b4 = *reinterpret_cast<const float4*>(Xt + c * N4 + j0);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float a = dUs[(i0 + r) * dp + c];
          acc[r][0] = fmaf(a, b4.x, acc[r][0]); acc[r][1] = fmaf(a, b4.y, acc[r][1]);
          acc[r][2] = fmaf(a, b4.z, acc[r][2]); acc[r][3] = fmaf(a, b4.w, acc[r][3]);
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (i0 + r < n && j0 + e < n) dab[(long long)(i0 + r) * N + j0 + e] += acc[r][e];
    }
  }
}

static bool small_enabled() {
  static int en = -1;
  if (en < 0) { const char* e = getenv("GP_NO_SMALL_GCN"); en = (e != nullptr && atoi(e) != 0) ? 0 : 1; }
  return en != 0;
}

// true if the fused per-graph kernels take this shape (forward and backward must agree)
bool small_gcn_eligible(int B, int N, int din, int dout, int add_self) {
  if (!small_enabled() || add_self || N > 128 || din > 128 || dout > 128 || B > 65535) return false;
  if (small_fwd_floats(N, din, dout) * 4 > (size_t)kSmallMaxSmem) return false;
  if (small_bwd_floats(N, din, dout, 1, 1) * 4 > (size_t)kSmallMaxSmem) return false;
  return true;
}

int small_gcn_fwd(const float* x, long long ldx, const float* adj, const float* w, const float* bias, const int32_t* nb,
                  int B, int N, int din, int dout, int normalize, float* u, float* y, long long ldy, float* rnorm,
                  cudaStream_t st) {
  GP_CONFIG_ONCE(
      GP_CUDA(cudaFuncSetAttribute(gconv_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallMaxSmem));
      GP_CUDA(cudaFuncSetAttribute(gconv_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallMaxSmem)));
  SmallFwd p;
  p.x = x; p.ldx = ldx; p.adj = adj; p.w = w; p.bias = bias; p.nb = nb;
  p.B = B; p.N = N; p.din = din; p.dout = dout; p.normalize = normalize;
  p.u = u; p.y = y; p.ldy = ldy; p.rnorm = rnorm;
  gconv_small_fwd_kernel<<<B, 256, small_fwd_floats(N, din, dout) * 4, st>>>(p);
  GP_LAUNCHED();
  return GP_OK;
}

// ws: B * (din*dout + dout) partial floats followed by the column-sum scratch (256 * (din*dout + dout) floats)
long long small_gcn_bwd_ws(int B, int din, int dout) {
  const long long w = (long long)din * dout + dout;
  return (long long)B * w + 256 * w;
}

int small_gcn_bwd(const float* dv, const float* u, const float* x, long long ldx, const float* adj, const float* w,
                  const int32_t* nb, int B, int N, int din, int dout, float* dw, float* db, float* dx, float* dadj,
                  float* ws, cudaStream_t st) {
  GP_CONFIG_ONCE(
      GP_CUDA(cudaFuncSetAttribute(gconv_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallMaxSmem));
      GP_CUDA(cudaFuncSetAttribute(gconv_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallMaxSmem)));
  SmallBwd p;
  p.dv = dv; p.u = u; p.x = x; p.ldx = ldx; p.adj = adj; p.w = w; p.nb = nb;
  p.B = B; p.N = N; p.din = din; p.dout = dout; p.part = ws; p.dx = dx; p.dadj = dadj;
  const int need_dx = dx != nullptr, need_da = dadj != nullptr;
  gconv_small_bwd_kernel<<<B, 256, small_bwd_floats(N, din, dout, need_dx, need_da) * 4, st>>>(p, need_dx, need_da);
  GP_LAUNCHED();
  const int wdt = din * dout + dout;
  float* tmp = ws + (long long)B * wdt;
  // reduce the per-graph partials over the batch (two-stage, deterministic); dW and db are contiguous slices
  if (db == dw + (long long)din * dout) {                // the caller laid dW | db out contiguously: one reduction
    GP_TRY(colsum(ws, B, wdt, wdt, dw, 0, tmp, st));
  } else {
    GP_TRY(colsum(ws, B, din * dout, wdt, dw, 0, tmp, st));
    if (db != nullptr) GP_TRY(colsum(ws + (long long)din * dout, B, dout, wdt, db, 0, tmp, st));
  }
  return GP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Per-graph fused POOLING for small graphs (encoders.py:1278-1279): X' = S^T Z, T = S^T A, A' = T S as ONE kernel,
// one CTA per graph, with S, Z, the real n_b x n_b adjacency block and the intermediate T = S^T A all in shared
// memory (T is chained straight into the second product; it is also written out because the backward API takes
// it).  Backward (no dA: level 0, or any level when the caller passes dadj == NULL):
//     dZ (+)= S dX'        dS (+)= Z dX'^T + T^T dA' + A (S dA'^T)
// again one kernel, every operand staged once.  Replaces 3 (forward) and 5 (backward) batched-GEMM launches that are
// latency-bound at these sizes.  S rows beyond n_b are zero by construction (masked softmax), so only the real rows
// are staged; pad rows / columns of the outputs are written as zeros (or left alone when accumulating).
// ---------------------------------------------------------------------------------------------------------
struct SmallPool {
  const float* s; const float* z; long long ldz; const float* adj; const int32_t* nb;
  int B, N, K, F;
  float* xp; float* t; float* ap;                                     // forward outputs
  const float* dxp; const float* dap; const float* tin;               // backward inputs
  float* dz; long long lddz; int acc_dz; float* ds; int acc_ds;       // backward outputs
};

__host__ __device__ inline size_t small_pool_fwd_floats(int N, int K, int F) {
  const size_t n4 = r4i(N), k4 = r4i(K), f4 = r4i(F);
  const size_t wide = f4 > k4 ? f4 : k4;
  return n4 * k4 + k4 * n4 + n4 * f4 + n4 * n4 + k4 * n4 + k4 * wide;
}
__host__ __device__ inline size_t small_pool_bwd_floats(int N, int K, int F) {
  const size_t n4 = r4i(N), k4 = r4i(K), f4 = r4i(F);
  return 4 * n4 * k4 + n4 * f4 + n4 * n4 + 2 * k4 * f4 + 2 * k4 * k4;
}

__global__ void __launch_bounds__(256) pool_small_fwd_kernel(const SmallPool p) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int N = p.N, K = p.K, F = p.F;
  const int n = p.nb != nullptr ? min(max(p.nb[b], 0), N) : N;
  const int n4 = r4i(n), k4 = r4i(K), f4 = r4i(F);
  const int N4 = r4i(N);
  const int wide = f4 > k4 ? f4 : k4;
  float* Ss = sm;                                   // [N4][k4]  S (real rows)
  float* St = Ss + (size_t)N4 * k4;                 // [k4][N4]  S^T
  float* Zs = St + (size_t)k4 * N4;                 // [N4][f4]
  float* As = Zs + (size_t)N4 * f4;                 // [N4][N4]  real adjacency block
  float* Ts = As + (size_t)N4 * N4;                 // [k4][N4]  T = S^T A  (stays on chip for A' = T S)
  float* Cs = Ts + (size_t)k4 * N4;                 // [k4][wide] output staging
  const float* sb = p.s + (long long)b * N * K;
  stage_rows(Ss, k4, sb, K, n, K, n4, k4);
  stage_rows_t(St, n4, sb, K, n, K, n4, k4);
  stage_rows(Zs, f4, p.z + (long long)b * N * p.ldz, p.ldz, n, F, n4, f4);
  stage_rows(As, n4, p.adj + (long long)b * N * N, N, n, n, n4, n4);
  cpa_wait_all();
  __syncthreads();
  smem_gemm(St, n4, Zs, f4, K, F, n, Cs, wide, nullptr);            // X' = S^T Z
  __syncthreads();
  float* xpb = p.xp + (long long)b * K * F;
  for (int e = tid; e < K * F; e += 256) { const int k = e / F, c = e - k * F; xpb[e] = Cs[k * wide + c]; }
  smem_gemm(St, n4, As, n4, K, n, n, Ts, n4, nullptr);              // T = S^T A
  __syncthreads();
  float* tb = p.t + (long long)b * K * N;
  for (int e = tid; e < K * N; e += 256) { const int k = e / N, j = e - k * N; tb[e] = j < n ? Ts[k * n4 + j] : 0.f; }
  smem_gemm(Ts, n4, Ss, k4, K, K, n, Cs, wide, nullptr);            // A' = T S   (Cs: X' was copied out above)
  __syncthreads();
  float* apb = p.ap + (long long)b * K * K;
  for (int e = tid; e < K * K; e += 256) { const int k = e / K, c = e - k * K; apb[e] = Cs[k * wide + c]; }
}

__global__ void __launch_bounds__(256) pool_small_bwd_kernel(const SmallPool p) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int N = p.N, K = p.K, F = p.F;
  const int n = p.nb != nullptr ? min(max(p.nb[b], 0), N) : N;
  const int n4 = r4i(n), k4 = r4i(K), f4 = r4i(F);
  const int N4 = r4i(N);
  float* Ss = sm;                                   // [N4][k4]
  float* Tt = Ss + (size_t)N4 * k4;                 // [N4][k4]  T^T
  float* Wv = Tt + (size_t)N4 * k4;                 // [N4][k4]  W = S dA'^T
  float* dSs = Wv + (size_t)N4 * k4;                // [N4][k4]
  float* Zs = dSs + (size_t)N4 * k4;                // [N4][f4]
  float* As = Zs + (size_t)N4 * f4;                 // [N4][N4]
  float* dXs = As + (size_t)N4 * N4;                // [k4][f4]  dX'
  float* dXt = dXs + (size_t)k4 * f4;               // [f4][k4]  dX'^T
  float* dAs = dXt + (size_t)k4 * f4;               // [k4][k4]  dA'
  float* dAt = dAs + (size_t)k4 * k4;               // [k4][k4]  dA'^T
  const float* dxb = p.dxp + (long long)b * K * F;
  const float* dab = p.dap + (long long)b * K * K;
  stage_rows(Ss, k4, p.s + (long long)b * N * K, K, n, K, n4, k4);
  stage_rows_t(Tt, k4, p.tin + (long long)b * K * N, N, K, n, k4, n4);
  stage_rows(Zs, f4, p.z + (long long)b * N * p.ldz, p.ldz, n, F, n4, f4);
  stage_rows(As, n4, p.adj + (long long)b * N * N, N, n, n, n4, n4);
  stage_rows(dXs, f4, dxb, F, K, F, k4, f4);
  stage_rows_t(dXt, k4, dxb, F, K, F, k4, f4);
  stage_rows(dAs, k4, dab, K, K, K, k4, k4);
  stage_rows_t(dAt, k4, dab, K, K, K, k4, k4);
  cpa_wait_all();
  __syncthreads();
  float* dzb = p.dz + (long long)b * N * p.lddz;
  if (p.acc_dz) smem_gemm_to_global<true>(Ss, k4, dXs, f4, n, F, K, dzb, p.lddz);      // dZ (+)= S dX' (real rows)
  else          smem_gemm_to_global<false>(Ss, k4, dXs, f4, n, F, K, dzb, p.lddz);
  if (!p.acc_dz)                                                     // pad rows of dZ are zero
    for (int e = tid; e < (N - n) * F; e += 256) { const int i = n + e / F, c = e % F; dzb[(long long)i * p.lddz + c] = 0.f; }
  smem_gemm(Zs, f4, dXt, k4, n, K, F, dSs, k4, nullptr);            // dS = Z dX'^T
  smem_gemm(Ss, k4, dAt, k4, n, K, K, Wv, k4, nullptr);             // W = S dA'^T
  __syncthreads();
  smem_gemm<true>(Tt, k4, dAs, k4, n, K, K, dSs, k4, nullptr);      // dS += T^T dA'
  __syncthreads();
  smem_gemm<true>(As, n4, Wv, k4, n, K, n, dSs, k4, nullptr);       // dS += A W
  __syncthreads();
  float* dsb = p.ds + (long long)b * N * K;
  for (int e = tid; e < N * K; e += 256) {
    const int i = e / K, c = e - i * K;
    const float v = i < n ? dSs[i * k4 + c] : 0.f;
    if (p.acc_ds) { if (i < n) dsb[e] += v; } else dsb[e] = v;
  }
}

static bool small_pool_enabled() {
  static int en = -1;
  if (en < 0) { const char* e = getenv("GP_NO_SMALL_POOL"); en = (e != nullptr && atoi(e) != 0) ? 0 : 1; }
  return en != 0;
}

bool small_pool_eligible(int B, int N, int K, int F) {
  if (!small_pool_enabled() || N > 128 || K > 64 || F > 512 || B > 2147483647 / 2) return false;
  return small_pool_fwd_floats(N, K, F) * 4 <= (size_t)kSmallMaxSmem && small_pool_bwd_floats(N, K, F) * 4 <= (size_t)kSmallMaxSmem;
}

static int small_pool_cfg() {
  GP_CONFIG_ONCE(
      GP_CUDA(cudaFuncSetAttribute(pool_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallMaxSmem));
      GP_CUDA(cudaFuncSetAttribute(pool_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallMaxSmem)));
  return GP_OK;
}

int small_pool_fwd(const float* s, const float* z, long long ldz, const float* adj, const int32_t* nb, int B, int N,
                   int K, int F, float* xp, float* t, float* ap, cudaStream_t st) {
  GP_TRY(small_pool_cfg());
  SmallPool p = {};
  p.s = s; p.z = z; p.ldz = ldz; p.adj = adj; p.nb = nb; p.B = B; p.N = N; p.K = K; p.F = F;
  p.xp = xp; p.t = t; p.ap = ap;
  pool_small_fwd_kernel<<<B, 256, small_pool_fwd_floats(N, K, F) * 4, st>>>(p);
  GP_LAUNCHED();
  return GP_OK;
}

int small_pool_bwd(const float* dxp, const float* dap, const float* s, const float* z, long long ldz, const float* adj,
                   const float* t, const int32_t* nb, int B, int N, int K, int F, float* dz, long long lddz, int acc_dz,
                   float* ds, int acc_ds, cudaStream_t st) {
  GP_TRY(small_pool_cfg());
  SmallPool p = {};
  p.s = s; p.z = z; p.ldz = ldz; p.adj = adj; p.nb = nb; p.B = B; p.N = N; p.K = K; p.F = F;
  p.dxp = dxp; p.dap = dap; p.tin = t; p.dz = dz; p.lddz = lddz; p.acc_dz = acc_dz; p.ds = ds; p.acc_ds = acc_ds;
  pool_small_bwd_kernel<<<B, 256, small_pool_bwd_floats(N, K, F) * 4, st>>>(p);
  GP_LAUNCHED();
  return GP_OK;
}

}  // namespace gp

extern "C" long long gp_graphconv_bwd_ws(int B, int N, int din, int dout, int add_self) {
  long long f = 256LL * dout;                                        // column sums of the generic path
  if (gp::small_gcn_eligible(B, N, din, dout, add_self)) {
    const long long s = gp::small_gcn_bwd_ws(B, din, dout);
    if (s > f) f = s;
  }
  return f;
}
