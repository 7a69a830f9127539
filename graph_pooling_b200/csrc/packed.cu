// PACKED schedule for ENZYMES-sized graphs (include/gp_b200.h, "PACKED schedule"; blueprint with every formula:
// tests/packed_blueprint.py, pinned against the oracle on the CPU).
//
// Reference semantics (file:line under /root/reference):
//   GraphConv  y = normalize((A x) W + b)                      encoders.py:315-328
//   ReLU + BatchNorm1d(N) per node index, concat               encoders.py:1054-1081, 1048-1052
//   S = softmax(Linear(za)) * mask, X' = S^T Z, A' = S^T A S   encoders.py:1269-1279
//   max readout                                                encoders.py:1257, 1287
//   link loss                                                  encoders.py:1311-1331
//
// Why another schedule: at N = 100, d = 30 the padded [B, N, d] layout makes every per-node-index BatchNorm access a
// 120-byte fragment at a 12 KB pitch, 68 % of the rows are pad rows, and a step is ~105 dependent launches
// (profiles/r1x_launches_cfg1_b4096.md).  Here only real rows exist, a phase is ONE launch for all graphs, and every
// graph-structured product stays inside a CTA's shared memory.
#include "common.cuh"

namespace gp {
namespace pk {

constexpr int kMaxN = 128;
constexpr float kEpsNorm = 1e-12f;
constexpr double kEpsBn = 1e-5;
constexpr float kEpsLink = 1e-7f;
constexpr int kThreads = 256;

__host__ __device__ __forceinline__ int r4(int v) { return (v + 3) & ~3; }

// ------------------------------------------------------------------------------------------------------------------
// tiling helpers
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int t_row0(const gp_pk_tiling& t, int g) { return t.rowptr ? t.rowptr[g] : g * t.nfix; }
__device__ __forceinline__ int t_count(const gp_pk_tiling& t) {
  return t.ntiles ? *t.ntiles : (t.B + t.gpt - 1) / t.gpt;
}
__device__ __forceinline__ void t_graphs(const gp_pk_tiling& t, int tile, int& g0, int& g1) {
  if (t.tile_g0) {
    g0 = t.tile_g0[tile];
    g1 = t.tile_g0[tile + 1];
  } else {
    g0 = tile * t.gpt;
    g1 = min(t.B, g0 + t.gpt);
  }
}

// per-row maps of a tile: s_gs[i] = local row of the first row of i's graph, s_gid[i] = its graph
__device__ __forceinline__ void row_map(const gp_pk_tiling& t, int g0, int g1, int r0, int nt, int* s_gs, int* s_gid) {
  for (int i = threadIdx.x; i < nt; i += blockDim.x) {
    int g;
    if (t.rowptr) {
      int lo = g0, hi = g1;                       // last g in [g0, g1) with rowptr[g] <= r0 + i
      const int r = r0 + i;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (t.rowptr[mid] <= r) lo = mid; else hi = mid;
      }
      g = lo;
      // graphs with zero rows share a start: take the LAST graph starting at or before r that is non-empty
      s_gs[i] = t.rowptr[g] - r0;
    } else {
      g = g0 + i / t.nfix;
      s_gs[i] = (g - g0) * t.nfix;
    }
    s_gid[i] = g;
  }
}

struct Carve {
  float* p;
  __device__ float* take(int nfloats) { float* r = p; p += (nfloats + 3) & ~3; return r; }
};
struct Count {
  size_t n = 0;
  void take(int nfloats) { n += (size_t)((nfloats + 3) & ~3); }
};

// ------------------------------------------------------------------------------------------------------------------
// ReLU + BatchNorm statistics of one source: mean / 1/std per node index from the sums over the real rows plus the
// analytic pad rows (cnt_pad[n] copies of relu(normalize(bias))).   BatchNorm1d(N) over (B, d), biased variance.
// ------------------------------------------------------------------------------------------------------------------
__device__ void bn_stats(const gp_pk_src& src, const float* __restrict__ cnt_pad, int Nn, int B, float* s_mean,
                         float* s_istd) {
  if (src.sums == nullptr) return;
  double p1 = 0.0, p2 = 0.0;
  if (src.bias != nullptr && cnt_pad != nullptr) {        // every thread recomputes the d-term constants (d <= 128)
    double nn = 0.0;
    for (int c = 0; c < src.d; ++c) nn += (double)src.bias[c] * (double)src.bias[c];
    const float nrm = fmaxf(sqrtf((float)nn), kEpsNorm);
    for (int c = 0; c < src.d; ++c) {
      const float y = fmaxf(src.bias[c] / nrm, 0.f);
      p1 += y;
      p2 += (double)y * y;
    }
  }
  const double cnt = (double)B * (double)src.d;
  for (int n = threadIdx.x; n < Nn; n += blockDim.x) {
    const double cp = cnt_pad ? (double)cnt_pad[n] : 0.0;
    const double m = (src.sums[n] + cp * p1) / cnt;
    const double v = (src.sums[Nn + n] + cp * p2) / cnt - m * m;
    s_mean[n] = (float)m;
    s_istd[n] = (float)(1.0 / sqrt(fmax(v, 0.0) + kEpsBn));
  }
}

// rows [r0, r0+nt) of a source -> dst[i*ldd + coff + c], c < d, ReLU + BatchNorm applied when the source has sums.
// Columns [d, dz) are written as zeros (dz = d rounded up when the slot is the last one of the destination).
__device__ void load_rows(const gp_pk_src& src, int N, int r0, int nt, const int* s_gs, const int* s_gid,
                          const float* s_mean, const float* s_istd, float* dst, int ldd, int coff, int dz) {
  const int d = src.d;
  const bool bn = src.sums != nullptr;
  for (int idx = threadIdx.x; idx < nt * dz; idx += blockDim.x) {
    const int i = idx / dz, c = idx - i * dz;
    float v = 0.f;
    if (c < d) {
      const int ni = i - s_gs[i];
      const long long row = src.padded ? ((long long)s_gid[i] * N + ni) : (long long)(r0 + i);
      v = src.y[row * src.ld + c];
      if (bn) v = (fmaxf(v, 0.f) - s_mean[ni]) * s_istd[ni];
    }
    dst[i * ldd + coff + c] = v;
  }
}

// dst[i][c] = sum_e a(i, e) * src[gs(i) + col_e][c]   (one warp per row; c < ld)
__device__ void gather(const gp_pk_adj& a, int nfix, int r0, int nt, const int* s_gs, const int* s_gid,
                       const float* src, int ld, float* dst) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = w; i < nt; i += nw) {
    const int gs = s_gs[i], ni = i - gs;
    int start = 0, deg = nfix;
    const float* dn = nullptr;
    if (a.info) {
      const int2 inf = reinterpret_cast<const int2*>(a.info)[r0 + i];
      start = inf.x;
      deg = inf.y;
    } else {
      dn = a.dense + (long long)s_gid[i] * nfix * nfix;
    }
    for (int c = lane; c < ld; c += 32) {
      float acc = 0.f;
      for (int e = 0; e < deg; ++e) {
        int col;
        float val;
        if (a.info) {
          const int2 en = reinterpret_cast<const int2*>(a.entries)[start + e];
          col = en.x;
          val = __int_as_float(en.y);
        } else {
          col = e;
          val = a.transposed ? dn[e * nfix + ni] : dn[ni * nfix + e];
        }
        acc = fmaf(val, src[(gs + col) * ld + c], acc);
      }
      dst[i * ld + c] = acc;
    }
  }
}

// C[m][n] = sum_k A[m*lda + k] * B[k*ldb + n]   m < M, n < 4*N4, k < 4*K4 (operands zero-padded to the 4-multiples)
template <class Epi>
__device__ __forceinline__ void gemm_nn(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
                                        int M, int N4, int K4, Epi epi) {
  const int MB = (M + 3) >> 2;
  for (int blk = threadIdx.x; blk < MB * N4; blk += blockDim.x) {
    const int i = blk / N4, j = blk - i * N4;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    const int mrem = M - 4 * i;
    const float* a0 = A + (size_t)(4 * i) * lda;
    const float* a1 = a0 + (mrem > 1 ? lda : 0);
    const float* a2 = a0 + (mrem > 2 ? 2 * lda : 0);
    const float* a3 = a0 + (mrem > 3 ? 3 * lda : 0);
    const float* b = Bm + 4 * j;
    for (int k4 = 0; k4 < K4; ++k4) {
      const float4 x0 = *reinterpret_cast<const float4*>(a0 + 4 * k4);
      const float4 x1 = *reinterpret_cast<const float4*>(a1 + 4 * k4);
      const float4 x2 = *reinterpret_cast<const float4*>(a2 + 4 * k4);
      const float4 x3 = *reinterpret_cast<const float4*>(a3 + 4 * k4);
      const float xs[4][4] = {{x0.x, x0.y, x0.z, x0.w}, {x1.x, x1.y, x1.z, x1.w}, {x2.x, x2.y, x2.z, x2.w},
                              {x3.x, x3.y, x3.z, x3.w}};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 wv = *reinterpret_cast<const float4*>(b + (size_t)(4 * k4 + kk) * ldb);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          acc[r][0] = fmaf(xs[r][kk], wv.x, acc[r][0]);
          acc[r][1] = fmaf(xs[r][kk], wv.y, acc[r][1]);
          acc[r][2] = fmaf(xs[r][kk], wv.z, acc[r][2]);
          acc[r][3] = fmaf(xs[r][kk], wv.w, acc[r][3]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r < mrem) {
#pragma unroll
        for (int c = 0; c < 4; ++c) epi(4 * i + r, 4 * j + c, acc[r][c]);
      }
  }
}

// Cs[grp][m*ldc + n] += sum_{r = rb + grp, step ngroups, < re} A[r*lda + m] * B[r*ldb + n]   m < 4*M4, n < 4*N4.
// Every (group, 4x4 block) item is owned by one fixed thread: plain read-modify-write, deterministic.
__device__ __forceinline__ void gemm_tn_acc(const float* __restrict__ A, int lda, const float* __restrict__ Bm,
                                            int ldb, int M4, int N4, int rb, int re, float* Cs, int ldc, int ngroups) {
  const int nblk = M4 * N4;
  const int gstride = 4 * M4 * ldc;
  for (int item = threadIdx.x; item < nblk * ngroups; item += blockDim.x) {
    const int grp = item / nblk, blk = item - grp * nblk;
    const int i = blk / N4, j = blk - i * N4;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    for (int r = rb + grp; r < re; r += ngroups) {
      const float4 a = *reinterpret_cast<const float4*>(A + (size_t)r * lda + 4 * i);
      const float4 b = *reinterpret_cast<const float4*>(Bm + (size_t)r * ldb + 4 * j);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        acc[rr][0] = fmaf(av[rr], b.x, acc[rr][0]);
        acc[rr][1] = fmaf(av[rr], b.y, acc[rr][1]);
        acc[rr][2] = fmaf(av[rr], b.z, acc[rr][2]);
        acc[rr][3] = fmaf(av[rr], b.w, acc[rr][3]);
      }
    }
    float* c = Cs + (size_t)grp * gstride + (size_t)(4 * i) * ldc + 4 * j;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr)
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) c[rr * ldc + cc] += acc[rr][cc];
  }
}

__device__ __forceinline__ int tn_groups(int M4, int N4, int cap_floats, int ldc) {
  const int nblk = M4 * N4;
  int g = blockDim.x / max(nblk, 1);
  g = max(1, min(g, 4));
  while (g > 1 && g * 4 * M4 * ldc > cap_floats) --g;
  return g;
}

// value of an upstream-gradient source at (local row i, column c)
__device__ __forceinline__ float grad_at(const gp_pk_grad& g, int r0, int i, int gid, int ni, int c) {
  float v = 0.f;
  if (g.dense) v = g.dense[(long long)(r0 + i) * g.ld + g.coff + c];
  if (g.dout) {
    const long long o = (long long)gid * g.ldo + g.ooff + c;
    if (g.arg[o] == ni) v += g.dout[o];
  }
  return v;
}

// ------------------------------------------------------------------------------------------------------------------
// prepare: scan of the node counts, pad counts per node index, the two tilings
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
prepare_kernel(const int32_t* __restrict__ nb, int B, int N, int w1, int w2, int32_t* __restrict__ rowptr,
               float* __restrict__ cnt_pad, int32_t* __restrict__ tiles1, int32_t* __restrict__ tiles2,
               int32_t* __restrict__ meta) {
  __shared__ int s_part[1024];
  __shared__ int s_hist[kMaxN + 2];
  const int tid = threadIdx.x, nth = blockDim.x;
  for (int i = tid; i <= N + 1; i += nth) s_hist[i] = 0;
  __syncthreads();
  const int per = (B + nth - 1) / nth;
  const int b0 = min(B, tid * per), b1 = min(B, b0 + per);
  int loc = 0;
  for (int b = b0; b < b1; ++b) {
    const int n = nb ? min(max(nb[b], 0), N) : N;
    loc += n;
    atomicAdd(&s_hist[n], 1);
  }
  s_part[tid] = loc;
  __syncthreads();
  for (int o = 1; o < nth; o <<= 1) {          // Hillis-Steele inclusive scan
    const int v = tid >= o ? s_part[tid - o] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  int run = tid ? s_part[tid - 1] : 0;
  for (int b = b0; b < b1; ++b) {
    rowptr[b] = run;
    run += nb ? min(max(nb[b], 0), N) : N;
  }
  const int R = s_part[nth - 1];
  if (tid == 0) rowptr[B] = R;
  // cnt_pad[n] = #graphs with n_b <= n
  if (tid == 0) {
    int c = 0;
    for (int n = 0; n < N; ++n) {
      c += s_hist[n];
      cnt_pad[n] = (float)c;
    }
  }
  __threadfence_block();
  __syncthreads();
  // tile t = graphs whose first row is in [t*w, (t+1)*w): tiles[t] = lower_bound(rowptr[0..B), t*w)
  for (int pass = 0; pass < 2; ++pass) {
    const int w = pass ? w2 : w1;
    int32_t* tiles = pass ? tiles2 : tiles1;
    const int nt = (R + w - 1) / w;
    for (int t = tid; t <= nt; t += nth) {
      if (t == nt) { tiles[t] = B; continue; }
      const int target = t * w;
      int lo = 0, hi = B;                       // first g with rowptr[g] >= target
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (rowptr[mid] < target) lo = mid + 1; else hi = mid;
      }
      tiles[t] = lo;
    }
    if (tid == 0) meta[1 + pass] = nt;
  }
  if (tid == 0) { meta[0] = R; meta[3] = 0; }
}

// ------------------------------------------------------------------------------------------------------------------
// neighbour lists of the dense level-0 adjacency (one CTA per graph at a time)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
build_lists_kernel(const float* __restrict__ adj, const int32_t* __restrict__ nb, const int32_t* __restrict__ rowptr,
                   int B, int N, int2* __restrict__ info_out, int2* __restrict__ ent_out, int2* __restrict__ info_in,
                   int2* __restrict__ ent_in, int32_t* __restrict__ cursor, long long capacity) {
  extern __shared__ __align__(16) float sm[];
  float* s_a = sm;                                   // [n][n+1]
  int* s_do = reinterpret_cast<int*>(sm + N * (N + 1));   // out degree / offset [N+1]
  int* s_di = s_do + N + 1;                          // in degree / offset
  __shared__ int s_base;
  const int tid = threadIdx.x;
  for (int g = blockIdx.x; g < B; g += gridDim.x) {
    const int n = nb ? min(max(nb[g], 0), N) : N;
    const int r0 = rowptr[g];
    const float* ag = adj + (long long)g * N * N;
    const int ld = n + 1;
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += blockDim.x) {
      const int i = idx / n, j = idx - i * n;
      s_a[i * ld + j] = ag[(long long)i * N + j];
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
      int co = 0, ci = 0;
      for (int j = 0; j < n; ++j) {
        co += s_a[i * ld + j] != 0.f;
        ci += s_a[j * ld + i] != 0.f;
      }
      s_do[i] = co;
      s_di[i] = ci;
    }
    __syncthreads();
    if (tid == 0) {
      int ro = 0, ri = 0;
      for (int i = 0; i < n; ++i) {
        const int a = s_do[i], b = s_di[i];
        s_do[i] = ro;
        s_di[i] = ri;
        ro += a;
        ri += b;
      }
      s_do[n] = ro;
      s_di[n] = ri;
      s_base = ro > 0 ? atomicAdd(cursor, ro) : 0;     // both lists hold the same number of entries
    }
    __syncthreads();
    const long long base = s_base;
    const bool ok = base + s_do[n] <= capacity;        // cannot fail when capacity >= sum n_b^2
    for (int i = tid; i < n; i += blockDim.x) {
      const int so = s_do[i], si = s_di[i];
      info_out[r0 + i] = make_int2((int)(base + so), ok ? s_do[i + 1] - so : 0);
      info_in[r0 + i] = make_int2((int)(base + si), ok ? s_di[i + 1] - si : 0);
      if (!ok) continue;
      int po = 0, pi = 0;
      for (int j = 0; j < n; ++j) {
        const float vo = s_a[i * ld + j];
        if (vo != 0.f) ent_out[base + so + po++] = make_int2(j, __float_as_int(vo));
        const float vi = s_a[j * ld + i];
        if (vi != 0.f) ent_in[base + si + pi++] = make_int2(j, __float_as_int(vi));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// GCN layer forward, one or two stacks in lock-step
// ------------------------------------------------------------------------------------------------------------------
struct FwdDims {
  int ldi[2], ldo[2], ldbuf;        // padded widths
};
__host__ __device__ inline FwdDims fwd_dims(const gp_pk_layer_fwd_args& p) {
  FwdDims d;
  d.ldbuf = 4;
  for (int s = 0; s < p.ns; ++s) {
    d.ldi[s] = r4(p.s[s].in.d);
    d.ldo[s] = r4(p.s[s].dout);
    d.ldbuf = d.ldbuf > d.ldi[s] ? d.ldbuf : d.ldi[s];
    d.ldbuf = d.ldbuf > d.ldo[s] ? d.ldbuf : d.ldo[s];
  }
  return d;
}
template <class C>
inline void fwd_carve(const gp_pk_layer_fwd_args& p, const FwdDims& d, C& c) {
  c.take(p.tl.max_rows * d.ldbuf);   // buf0
  c.take(p.tl.max_rows * d.ldbuf);   // buf1
  c.take(p.tl.max_rows);             // gs
  c.take(p.tl.max_rows);             // gid
  for (int s = 0; s < p.ns; ++s) {
    c.take(d.ldi[s] * d.ldo[s]);     // W
    c.take(d.ldo[s]);                // b
    c.take(2 * p.N);                 // mean / istd of the input
    c.take(2 * p.N);                 // partial sums
  }
}

__global__ void __launch_bounds__(kThreads)
layer_fwd_kernel(const gp_pk_layer_fwd_args p) {
  extern __shared__ __align__(16) float sm[];
  const FwdDims dm = fwd_dims(p);
  Carve cv{sm};
  float* buf0 = cv.take(p.tl.max_rows * dm.ldbuf);
  float* buf1 = cv.take(p.tl.max_rows * dm.ldbuf);
  int* s_gs = reinterpret_cast<int*>(cv.take(p.tl.max_rows));
  int* s_gid = reinterpret_cast<int*>(cv.take(p.tl.max_rows));
  float *s_w[2], *s_b[2], *s_bn[2], *s_st[2];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  const int N = p.N;
  for (int s = 0; s < p.ns; ++s) {
    s_w[s] = cv.take(dm.ldi[s] * dm.ldo[s]);
    s_b[s] = cv.take(dm.ldo[s]);
    s_bn[s] = cv.take(2 * N);
    s_st[s] = cv.take(2 * N);
    const gp_pk_stack_fwd& st = p.s[s];
    const int din = st.in.d, dout = st.dout, ldo = dm.ldo[s];
    for (int idx = tid; idx < dm.ldi[s] * ldo; idx += blockDim.x) {
      const int k = idx / ldo, n = idx - k * ldo;
      s_w[s][idx] = (k < din && n < dout) ? st.W[(long long)k * dout + n] : 0.f;
    }
    for (int n = tid; n < ldo; n += blockDim.x) s_b[s][n] = (st.b && n < dout) ? st.b[n] : 0.f;
    for (int n = tid; n < 2 * N; n += blockDim.x) s_st[s][n] = 0.f;
    bn_stats(st.in, p.cnt_pad, N, p.tl.B, s_bn[s], s_bn[s] + N);
  }
  __syncthreads();
  const int ntiles = t_count(p.tl);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int g0, g1;
    t_graphs(p.tl, tile, g0, g1);
    const int r0 = t_row0(p.tl, g0);
    const int nt = (g1 > g0 ? t_row0(p.tl, g1) : r0) - r0;
    if (nt <= 0) continue;
    __syncthreads();
    row_map(p.tl, g0, g1, r0, nt, s_gs, s_gid);
    __syncthreads();
    for (int s = 0; s < p.ns; ++s) {
      const gp_pk_stack_fwd& st = p.s[s];
      const int ldi = dm.ldi[s], ldo = dm.ldo[s], dout = st.dout;
      load_rows(st.in, N, r0, nt, s_gs, s_gid, s_bn[s], s_bn[s] + N, buf0, ldi, 0, ldi);
      __syncthreads();
      gather(p.adj, p.tl.nfix, r0, nt, s_gs, s_gid, buf0, ldi, buf1);
      __syncthreads();
      {
        float* v = buf0;
        const float* bb = s_b[s];
        gemm_nn(buf1, ldi, s_w[s], ldo, nt, ldo >> 2, ldi >> 2,
                [=](int m, int n, float a) { v[m * ldo + n] = a + bb[n]; });
      }
      __syncthreads();
      for (int i = wid; i < nt; i += nw) {
        const float* v = buf0 + i * ldo;
        float ss = 0.f;
        for (int c = lane; c < dout; c += 32) ss = fmaf(v[c], v[c], ss);
        ss = warp_sum(ss);
        const float r = fmaxf(sqrtf(ss), kEpsNorm);
        float a1 = 0.f, a2 = 0.f;
        float* yo = st.y + (long long)(r0 + i) * dout;
        for (int c = lane; c < dout; c += 32) {
          const float y = v[c] / r;
          yo[c] = y;
          const float q = fmaxf(y, 0.f);
          a1 += q;
          a2 = fmaf(q, q, a2);
        }
        if (lane == 0) st.rnorm[r0 + i] = r;
        if (st.sums_out) {
          a1 = warp_sum(a1);
          a2 = warp_sum(a2);
          if (lane == 0) {
            const int ni = i - s_gs[i];
            atomicAdd(&s_st[s][ni], a1);
            atomicAdd(&s_st[s][N + ni], a2);
          }
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  for (int s = 0; s < p.ns; ++s)
    if (p.s[s].sums_out)
      for (int n = tid; n < 2 * N; n += blockDim.x)
        if (s_st[s][n] != 0.f) atomicAdd(&p.s[s].sums_out[n], (double)s_st[s][n]);
}

// ------------------------------------------------------------------------------------------------------------------
// GCN layer backward
// ------------------------------------------------------------------------------------------------------------------
struct BwdDims {
  int ldi[2], ldo[2], ldbuf, dwcap;
};
__host__ __device__ inline BwdDims bwd_dims(const gp_pk_layer_bwd_args& p) {
  BwdDims d;
  d.ldbuf = 4;
  d.dwcap = 0;
  for (int s = 0; s < p.ns; ++s) {
    d.ldi[s] = r4(p.s[s].in.d);
    d.ldo[s] = r4(p.s[s].dout);
    d.ldbuf = d.ldbuf > d.ldi[s] ? d.ldbuf : d.ldi[s];
    d.ldbuf = d.ldbuf > d.ldo[s] ? d.ldbuf : d.ldo[s];
    const int one = d.ldi[s] * d.ldo[s];
    int groups = kThreads / ((d.ldi[s] >> 2) * (d.ldo[s] >> 2));
    groups = groups < 1 ? 1 : (groups > 4 ? 4 : groups);
    d.dwcap = d.dwcap > one * groups ? d.dwcap : one * groups;
  }
  return d;
}
template <class C>
inline void bwd_carve(const gp_pk_layer_bwd_args& p, const BwdDims& d, C& c) {
  for (int i = 0; i < 4; ++i) c.take(p.tl.max_rows * d.ldbuf);
  c.take(p.tl.max_rows);
  c.take(p.tl.max_rows);
  for (int s = 0; s < p.ns; ++s) {
    c.take(d.ldi[s] * d.ldo[s]);     // W   [din x ldo]
    c.take(d.ldo[s] * d.ldi[s]);     // W^T [dout x ldi]
    c.take(d.dwcap);                 // dW accumulators (groups)
    c.take(d.ldo[s]);                // db accumulator
    c.take(2 * p.N);                 // mean / istd of the input
    c.take(2 * p.N);                 // mean / istd of the output
    c.take(2 * p.N);                 // m1 / m2 of the output
    c.take(2 * p.N);                 // partial sums for msums_prev
  }
  c.take(p.N);                       // scratch (pad rows)
}

__global__ void __launch_bounds__(kThreads)
layer_bwd_kernel(const gp_pk_layer_bwd_args p) {
  extern __shared__ __align__(16) float sm[];
  const BwdDims dm = bwd_dims(p);
  Carve cv{sm};
  float* bG = cv.take(p.tl.max_rows * dm.ldbuf);    // gl -> dV
  float* bY = cv.take(p.tl.max_rows * dm.ldbuf);    // Y  -> dX
  float* bH = cv.take(p.tl.max_rows * dm.ldbuf);    // Hin
  float* bU = cv.take(p.tl.max_rows * dm.ldbuf);    // U  -> dU
  int* s_gs = reinterpret_cast<int*>(cv.take(p.tl.max_rows));
  int* s_gid = reinterpret_cast<int*>(cv.take(p.tl.max_rows));
  float *s_w[2], *s_wt[2], *s_dw[2], *s_db[2], *s_bni[2], *s_bno[2], *s_m[2], *s_mp[2];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  const int N = p.N, B = p.tl.B;
  int groups[2];
  for (int s = 0; s < p.ns; ++s) {
    const gp_pk_stack_bwd& st = p.s[s];
    const int din = st.in.d, dout = st.dout, ldi = dm.ldi[s], ldo = dm.ldo[s];
    s_w[s] = cv.take(ldi * ldo);
    s_wt[s] = cv.take(ldo * ldi);
    s_dw[s] = cv.take(dm.dwcap);
    s_db[s] = cv.take(ldo);
    s_bni[s] = cv.take(2 * N);
    s_bno[s] = cv.take(2 * N);
    s_m[s] = cv.take(2 * N);
    s_mp[s] = cv.take(2 * N);
    groups[s] = tn_groups(ldi >> 2, ldo >> 2, dm.dwcap, ldo);
    for (int idx = tid; idx < ldi * ldo; idx += blockDim.x) {
      const int k = idx / ldo, n = idx - k * ldo;
      s_w[s][idx] = (k < din && n < dout) ? st.W[(long long)k * dout + n] : 0.f;
    }
    for (int idx = tid; idx < ldo * ldi; idx += blockDim.x) {
      const int n = idx / ldi, k = idx - n * ldi;
      s_wt[s][idx] = (k < din && n < dout) ? st.W[(long long)k * dout + n] : 0.f;
    }
    for (int idx = tid; idx < dm.dwcap; idx += blockDim.x) s_dw[s][idx] = 0.f;
    for (int n = tid; n < ldo; n += blockDim.x) s_db[s][n] = 0.f;
    for (int n = tid; n < 2 * N; n += blockDim.x) s_mp[s][n] = 0.f;
    bn_stats(st.in, p.cnt_pad, N, B, s_bni[s], s_bni[s] + N);
    bn_stats(st.out, p.cnt_pad, N, B, s_bno[s], s_bno[s] + N);
    if (st.msums) {
      const double cnt = (double)B * (double)dout;
      for (int n = tid; n < 2 * N; n += blockDim.x) s_m[s][n] = (float)(st.msums[n] / cnt);
    }
  }
  float* s_scr = cv.take(N);
  __syncthreads();

  // pad rows of a BatchNorm'd layer with a bias: no upstream gradient, but the batch means reach them; cnt_pad[n]
  // copies of one vector per node index feed the bias gradient (packed_blueprint.stack_backward)
  if (blockIdx.x == 0) {
    for (int s = 0; s < p.ns; ++s) {
      const gp_pk_stack_bwd& st = p.s[s];
      if (!(st.out.sums && st.b && st.db && p.cnt_pad)) continue;
      const int dout = st.dout;
      float nn = 0.f;
      for (int c = 0; c < dout; ++c) nn = fmaf(st.b[c], st.b[c], nn);
      const float rp = fmaxf(sqrtf(nn), kEpsNorm);
      const float *mean = s_bno[s], *istd = s_bno[s] + N, *m1 = s_m[s], *m2 = s_m[s] + N;
      for (int n = tid; n < N; n += blockDim.x) {        // proj[n] = sum_c yp[c] * dYp[n][c]
        float pr = 0.f;
        for (int c = 0; c < dout; ++c) {
          const float yp = st.b[c] / rp;
          const float hp = (fmaxf(yp, 0.f) - mean[n]) * istd[n];
          const float dy = yp > 0.f ? (-m1[n] - hp * m2[n]) * istd[n] : 0.f;
          pr = fmaf(yp, dy, pr);
        }
        s_scr[n] = pr;
      }
      __syncthreads();
      for (int c = tid; c < dout; c += blockDim.x) {
        const float yp = st.b[c] / rp;
        float acc = 0.f;
        for (int n = 0; n < N; ++n) {
          const float cp = p.cnt_pad[n];
          if (cp == 0.f) continue;
          const float hp = (fmaxf(yp, 0.f) - mean[n]) * istd[n];
          const float dy = yp > 0.f ? (-m1[n] - hp * m2[n]) * istd[n] : 0.f;
          const float dv = rp > kEpsNorm ? (dy - yp * s_scr[n]) / rp : dy / kEpsNorm;
          acc = fmaf(cp, dv, acc);
        }
        s_db[s][c] += acc;
      }
      __syncthreads();
    }
  }

  const int ntiles = t_count(p.tl);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int g0, g1;
    t_graphs(p.tl, tile, g0, g1);
    const int r0 = t_row0(p.tl, g0);
    const int nt = (g1 > g0 ? t_row0(p.tl, g1) : r0) - r0;
    if (nt <= 0) continue;
    __syncthreads();
    row_map(p.tl, g0, g1, r0, nt, s_gs, s_gid);
    __syncthreads();
    for (int s = 0; s < p.ns; ++s) {
      const gp_pk_stack_bwd& st = p.s[s];
      const int din = st.in.d, dout = st.dout, ldi = dm.ldi[s], ldo = dm.ldo[s];
      // ---- dV = d normalize . d(ReLU + BatchNorm) . gl
      for (int idx = tid; idx < nt * ldo; idx += blockDim.x) {
        const int i = idx / ldo, c = idx - i * ldo;
        float g = 0.f, y = 0.f;
        if (c < dout) {
          g = grad_at(st.gl, r0, i, s_gid[i], i - s_gs[i], c);
          y = st.out.y[(long long)(r0 + i) * st.out.ld + c];
        }
        bG[idx] = g;
        bY[idx] = y;
      }
      __syncthreads();
      const bool bn = st.out.sums != nullptr;
      for (int i = wid; i < nt; i += nw) {
        const int ni = i - s_gs[i];
        const float r = st.rnorm[r0 + i];
        float mean = 0.f, istd = 1.f, m1 = 0.f, m2 = 0.f;
        if (bn) { mean = s_bno[s][ni]; istd = s_bno[s][N + ni]; m1 = s_m[s][ni]; m2 = s_m[s][N + ni]; }
        float* g = bG + i * ldo;
        const float* y = bY + i * ldo;
        float pr = 0.f;
        for (int c = lane; c < dout; c += 32) {
          float dy = g[c];
          if (bn) {
            const float h = (fmaxf(y[c], 0.f) - mean) * istd;
            dy = y[c] > 0.f ? (dy - m1 - h * m2) * istd : 0.f;
          }
          g[c] = dy;
          pr = fmaf(y[c], dy, pr);
        }
        pr = warp_sum(pr);
        for (int c = lane; c < dout; c += 32)
          g[c] = r > kEpsNorm ? (g[c] - y[c] * pr) / r : g[c] / kEpsNorm;
      }
      __syncthreads();
      if (st.db)
        for (int c = tid; c < dout; c += blockDim.x) {
          float acc = 0.f;
          for (int i = 0; i < nt; ++i) acc += bG[i * ldo + c];
          s_db[s][c] += acc;
        }
      // ---- U = A Hin (recomputed), dW += U^T dV
      load_rows(st.in, N, r0, nt, s_gs, s_gid, s_bni[s], s_bni[s] + N, bH, ldi, 0, ldi);
      __syncthreads();
      gather(p.adj, p.tl.nfix, r0, nt, s_gs, s_gid, bH, ldi, bU);
      __syncthreads();
      gemm_tn_acc(bU, ldi, bG, ldo, ldi >> 2, ldo >> 2, 0, nt, s_dw[s], ldo, groups[s]);
      __syncthreads();
      if (!st.need_dx) continue;
      // ---- dU = dV W^T, dA (dense level), dX = A^T dU
      {
        float* du = bU;
        gemm_nn(bG, ldo, s_wt[s], ldi, nt, ldi >> 2, ldo >> 2, [=](int m, int n, float a) { du[m * ldi + n] = a; });
      }
      __syncthreads();
      if (st.dadj) {
        const int nf = p.tl.nfix;
        for (int idx = tid; idx < nt * nf; idx += blockDim.x) {
          const int i = idx / nf, j = idx - i * nf;                 // dA[g][ni][j] = <dU[i], Hin[gs + j]>
          const float* a = bU + i * ldi;
          const float* h = bH + (s_gs[i] + j) * ldi;
          float acc = 0.f;
          for (int c = 0; c < din; ++c) acc = fmaf(a[c], h[c], acc);
          float* o = st.dadj + ((long long)s_gid[i] * nf + (i - s_gs[i])) * nf + j;
          *o = st.dadj_acc ? *o + acc : acc;
        }
      }
      gather(p.adj_in, p.tl.nfix, r0, nt, s_gs, s_gid, bU, ldi, bY);
      __syncthreads();
      for (int i = wid; i < nt; i += nw) {
        const int ni = i - s_gs[i], gid = s_gid[i];
        float a1 = 0.f, a2 = 0.f;
        for (int c = lane; c < din; c += 32) {
          const float g = bY[i * ldi + c] + grad_at(st.gz_prev, r0, i, gid, ni, c);
          st.gl_prev[(long long)(r0 + i) * din + c] = g;
          a1 += g;
          a2 = fmaf(g, bH[i * ldi + c], a2);
        }
        if (st.msums_prev) {
          a1 = warp_sum(a1);
          a2 = warp_sum(a2);
          if (lane == 0) {
            atomicAdd(&s_mp[s][ni], a1);
            atomicAdd(&s_mp[s][N + ni], a2);
          }
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  for (int s = 0; s < p.ns; ++s) {
    const gp_pk_stack_bwd& st = p.s[s];
    const int din = st.in.d, dout = st.dout, ldi = dm.ldi[s], ldo = dm.ldo[s];
    const int gstride = ldi * ldo;
    for (int idx = tid; idx < din * dout; idx += blockDim.x) {
      const int k = idx / dout, n = idx - k * dout;
      float acc = 0.f;
      for (int g = 0; g < groups[s]; ++g) acc += s_dw[s][g * gstride + k * ldo + n];
      if (acc != 0.f) atomicAdd(&st.dW[idx], acc);
    }
    if (st.db)
      for (int c = tid; c < dout; c += blockDim.x)
        if (s_db[s][c] != 0.f) atomicAdd(&st.db[c], s_db[s][c]);
    if (st.msums_prev)
      for (int n = tid; n < 2 * N; n += blockDim.x)
        if (s_mp[s][n] != 0.f) atomicAdd(&st.msums_prev[n], (double)s_mp[s][n]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// pooling forward / backward
// ------------------------------------------------------------------------------------------------------------------
struct PoolDims {
  int ldF, ldFa, ldK, Fa, F, K;
};
__host__ __device__ inline PoolDims pool_dims(const gp_pk_pool_args& p) {
  PoolDims d;
  d.F = p.z.F; d.Fa = p.za.F; d.K = p.K;
  d.ldF = r4(d.F); d.ldFa = r4(d.Fa); d.ldK = r4(d.K);
  return d;
}
template <class C>
inline void pool_carve(const gp_pk_pool_args& p, const PoolDims& d, bool bwd, C& c) {
  const int mr = p.tl.max_rows;
  c.take(mr * d.ldF);        // Z
  c.take(mr * d.ldFa);       // Za
  c.take(mr * d.ldK);        // S
  c.take(mr * d.ldK);        // A^T S   (fwd: T^T; bwd)
  c.take(mr);                // gs
  c.take(mr);                // gid
  c.take((p.z.L + p.za.L) * 2 * p.N);   // BatchNorm of every slot
  if (!bwd) {
    c.take(d.ldFa * d.ldK);  // Wp^T [Fa x ldK]
    c.take(d.ldK);           // bp
  } else {
    c.take(mr * d.ldK);      // A S
    c.take(mr * d.ldK);      // dS -> dT
    c.take(d.ldK * d.ldFa);  // Wp [K x ldFa]
    c.take(d.ldK * d.ldF);   // dX' of the current graph
    c.take(d.ldK * d.ldK);   // dA' of the current graph
    c.take(d.ldK * d.ldFa);  // dWp accumulator
    c.take(d.ldK);           // dbp accumulator
  }
}

__device__ void concat_stats(const gp_pk_concat& z, const float* cnt_pad, int N, int B, float* s_bn) {
  for (int l = 0; l < z.L; ++l) bn_stats(z.slot[l], cnt_pad, N, B, s_bn + l * 2 * N, s_bn + l * 2 * N + N);
}
__device__ void concat_load(const gp_pk_concat& z, int N, int r0, int nt, const int* s_gs, const int* s_gid,
                            const float* s_bn, float* dst, int ldd) {
  int off = 0;
  for (int l = 0; l < z.L; ++l) {
    const int d = z.slot[l].d;
    const int dz = (l == z.L - 1) ? ldd - off : d;
    load_rows(z.slot[l], N, r0, nt, s_gs, s_gid, s_bn + l * 2 * N, s_bn + l * 2 * N + N, dst, ldd, off, dz);
    off += d;
  }
}

__global__ void __launch_bounds__(kThreads)
pool_fwd_kernel(const gp_pk_pool_args p) {
  extern __shared__ __align__(16) float sm[];
  const PoolDims dm = pool_dims(p);
  const int mr = p.tl.max_rows, N = p.N, K = dm.K, F = dm.F, Fa = dm.Fa, ldF = dm.ldF, ldFa = dm.ldFa, ldK = dm.ldK;
  Carve cv{sm};
  float* bZ = cv.take(mr * ldF);
  float* bZa = cv.take(mr * ldFa);
  float* bS = cv.take(mr * ldK);
  float* bT = cv.take(mr * ldK);
  int* s_gs = reinterpret_cast<int*>(cv.take(mr));
  int* s_gid = reinterpret_cast<int*>(cv.take(mr));
  float* s_bn = cv.take((p.z.L + p.za.L) * 2 * N);
  float* s_bna = s_bn + p.z.L * 2 * N;
  float* s_wpt = cv.take(ldFa * ldK);
  float* s_bp = cv.take(ldK);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  for (int idx = tid; idx < ldFa * ldK; idx += blockDim.x) {
    const int f = idx / ldK, k = idx - f * ldK;
    s_wpt[idx] = (f < Fa && k < K) ? p.Wp[(long long)k * Fa + f] : 0.f;
  }
  for (int k = tid; k < ldK; k += blockDim.x) s_bp[k] = (p.bp && k < K) ? p.bp[k] : 0.f;
  concat_stats(p.z, p.cnt_pad, N, p.tl.B, s_bn);
  concat_stats(p.za, p.cnt_pad, N, p.tl.B, s_bna);
  __syncthreads();
  const int ntiles = t_count(p.tl);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int g0, g1;
    t_graphs(p.tl, tile, g0, g1);
    const int r0 = t_row0(p.tl, g0);
    const int nt = (g1 > g0 ? t_row0(p.tl, g1) : r0) - r0;
    __syncthreads();
    // graphs of the tile without rows still own their (all-zero) outputs
    if (nt > 0) row_map(p.tl, g0, g1, r0, nt, s_gs, s_gid);
    __syncthreads();
    if (nt > 0) {
      concat_load(p.z, N, r0, nt, s_gs, s_gid, s_bn, bZ, ldF);
      concat_load(p.za, N, r0, nt, s_gs, s_gid, s_bna, bZa, ldFa);
    }
    __syncthreads();
    if (nt > 0) {
      float* sp = bS;
      const float* bb = s_bp;
      gemm_nn(bZa, ldFa, s_wpt, ldK, nt, ldK >> 2, ldFa >> 2, [=](int m, int n, float a) { sp[m * ldK + n] = a + bb[n]; });
    }
    __syncthreads();
    for (int i = wid; i < nt; i += nw) {                       // softmax over the K clusters
      float* s = bS + i * ldK;
      float mx = -INFINITY;
      for (int k = lane; k < K; k += 32) mx = fmaxf(mx, s[k]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int k = lane; k < K; k += 32) {
        const float e = expf(s[k] - mx);
        s[k] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      float* so = p.S + ((long long)s_gid[i] * N + (i - s_gs[i])) * K;
      for (int k = lane; k < ldK; k += 32) {
        const float v = k < K ? s[k] / sum : 0.f;
        s[k] = v;
        if (k < K) so[k] = v;
      }
    }
    for (int g = g0; g < g1; ++g) {                            // S rows of pad nodes are zero (mask, :1275)
      const int n = t_row0(p.tl, g + 1) - t_row0(p.tl, g);
      float* so = p.S + ((long long)g * N + n) * K;
      for (int idx = tid; idx < (N - n) * K; idx += blockDim.x) so[idx] = 0.f;
    }
    __syncthreads();
    if (nt > 0) gather(p.adj_in, p.tl.nfix, r0, nt, s_gs, s_gid, bS, ldK, bT);    // (A^T S)[j][k] = T[k][j]
    // max readout over the graph's rows; with pad rows (zeros after the mask) a negative maximum loses to 0
    const int ng = g1 - g0;
    for (int idx = tid; idx < ng * F; idx += blockDim.x) {
      const int gl = idx / F, f = idx - gl * F, g = g0 + gl;
      const int rs = t_row0(p.tl, g) - r0, n = t_row0(p.tl, g + 1) - r0 - rs;
      float best = -INFINITY;
      int arg = -1;
      for (int i = 0; i < n; ++i) {
        const float v = bZ[(rs + i) * ldF + f];
        if (v > best) { best = v; arg = i; }
      }
      if (n < N && !(best >= 0.f)) { best = 0.f; arg = -1; }
      p.out[(long long)g * p.ldo + f] = best;
      p.arg[(long long)g * p.ldo + f] = arg;
    }
    __syncthreads();
    // X'[g] = S^T Z  and  A'[g] = (A^T S)^T S, 4x4 blocks over (graph, k, f)
    {
      const int K4 = ldK >> 2, F4 = ldF >> 2;
      const int per = K4 * (F4 + K4);
      for (int item = tid; item < ng * per; item += blockDim.x) {
        const int gl = item / per, blk = item - gl * per, g = g0 + gl;
        const int rs = t_row0(p.tl, g) - r0, re = t_row0(p.tl, g + 1) - r0;
        const bool isx = blk < K4 * F4;
        const int b2 = isx ? blk : blk - K4 * F4;
        const int nb4 = isx ? F4 : K4;
        const int i = b2 / nb4, j = b2 - i * nb4;
        const float* A = isx ? bS : bT;
        const float* Bm = isx ? bZ : bS;
        const int ldb = isx ? ldF : ldK;
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
        for (int r = rs; r < re; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(A + r * ldK + 4 * i);
          const float4 b = *reinterpret_cast<const float4*>(Bm + r * ldb + 4 * j);
          const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            acc[rr][0] = fmaf(av[rr], b.x, acc[rr][0]);
            acc[rr][1] = fmaf(av[rr], b.y, acc[rr][1]);
            acc[rr][2] = fmaf(av[rr], b.z, acc[rr][2]);
            acc[rr][3] = fmaf(av[rr], b.w, acc[rr][3]);
          }
        }
        const int ncol = isx ? F : K;
        float* o = isx ? p.xp + (long long)g * K * F : p.ap + (long long)g * K * K;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
            if (4 * i + rr < K && 4 * j + cc < ncol) o[(4 * i + rr) * ncol + 4 * j + cc] = acc[rr][cc];
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads)
pool_bwd_kernel(const gp_pk_pool_args p) {
  extern __shared__ __align__(16) float sm[];
  const PoolDims dm = pool_dims(p);
  const int mr = p.tl.max_rows, N = p.N, K = dm.K, F = dm.F, Fa = dm.Fa, ldF = dm.ldF, ldFa = dm.ldFa, ldK = dm.ldK;
  Carve cv{sm};
  float* bZ = cv.take(mr * ldF);
  float* bZa = cv.take(mr * ldFa);
  float* bS = cv.take(mr * ldK);
  float* bAtS = cv.take(mr * ldK);
  int* s_gs = reinterpret_cast<int*>(cv.take(mr));
  int* s_gid = reinterpret_cast<int*>(cv.take(mr));
  float* s_bn = cv.take((p.z.L + p.za.L) * 2 * N);
  float* s_bna = s_bn + p.z.L * 2 * N;
  float* bAS = cv.take(mr * ldK);
  float* bD = cv.take(mr * ldK);
  float* s_wp = cv.take(ldK * ldFa);
  float* s_dxp = cv.take(ldK * ldF);
  float* s_dap = cv.take(ldK * ldK);
  float* s_dwp = cv.take(ldK * ldFa);
  float* s_dbp = cv.take(ldK);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  for (int idx = tid; idx < ldK * ldFa; idx += blockDim.x) {
    const int k = idx / ldFa, f = idx - k * ldFa;
    s_wp[idx] = (k < K && f < Fa) ? p.Wp[(long long)k * Fa + f] : 0.f;
    s_dwp[idx] = 0.f;
  }
  for (int k = tid; k < ldK; k += blockDim.x) s_dbp[k] = 0.f;
  concat_stats(p.z, p.cnt_pad, N, p.tl.B, s_bn);
  concat_stats(p.za, p.cnt_pad, N, p.tl.B, s_bna);
  __syncthreads();
  const int ntiles = t_count(p.tl);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int g0, g1;
    t_graphs(p.tl, tile, g0, g1);
    const int r0 = t_row0(p.tl, g0);
    const int nt = (g1 > g0 ? t_row0(p.tl, g1) : r0) - r0;
    if (nt <= 0) continue;
    __syncthreads();
    row_map(p.tl, g0, g1, r0, nt, s_gs, s_gid);
    __syncthreads();
    concat_load(p.z, N, r0, nt, s_gs, s_gid, s_bn, bZ, ldF);
    concat_load(p.za, N, r0, nt, s_gs, s_gid, s_bna, bZa, ldFa);
    for (int idx = tid; idx < nt * ldK; idx += blockDim.x) {
      const int i = idx / ldK, k = idx - i * ldK;
      bS[idx] = k < K ? p.S[((long long)s_gid[i] * N + (i - s_gs[i])) * K + k] : 0.f;
    }
    __syncthreads();
    gather(p.adj, p.tl.nfix, r0, nt, s_gs, s_gid, bS, ldK, bAS);
    gather(p.adj_in, p.tl.nfix, r0, nt, s_gs, s_gid, bS, ldK, bAtS);
    for (int g = g0; g < g1; ++g) {
      const int rs = t_row0(p.tl, g) - r0, n = t_row0(p.tl, g + 1) - r0 - rs;
      if (n <= 0) continue;
      __syncthreads();
      for (int idx = tid; idx < ldK * ldF; idx += blockDim.x) {
        const int k = idx / ldF, f = idx - k * ldF;
        s_dxp[idx] = (k < K && f < F) ? p.dxp[((long long)g * K + k) * F + f] : 0.f;
      }
      for (int idx = tid; idx < ldK * ldK; idx += blockDim.x) {
        const int k = idx / ldK, k2 = idx - k * ldK;
        s_dap[idx] = (k < K && k2 < K) ? p.dap[((long long)g * K + k) * K + k2] : 0.f;
      }
      __syncthreads();
      // dS[i][k] = <Z[i], dX'[k]> + <AS[i], dA'[k]> + sum_k2 AtS[i][k2] dA'[k2][k] + dS_ext
      for (int idx = tid; idx < n * ldK; idx += blockDim.x) {
        const int il = idx / ldK, k = idx - il * ldK, i = rs + il;
        float acc = 0.f;
        if (k < K) {
          const float* z = bZ + i * ldF;
          const float* dx = s_dxp + k * ldF;
          for (int f = 0; f < ldF; f += 4) {
            const float4 a = *reinterpret_cast<const float4*>(z + f);
            const float4 b = *reinterpret_cast<const float4*>(dx + f);
            acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
          }
          const float* as = bAS + i * ldK;
          const float* ats = bAtS + i * ldK;
          for (int k2 = 0; k2 < K; ++k2) {
            acc = fmaf(as[k2], s_dap[k * ldK + k2], acc);
            acc = fmaf(ats[k2], s_dap[k2 * ldK + k], acc);
          }
          if (p.dS_ext) acc += p.dS_ext[((long long)g * N + il) * K + k];
        }
        bD[i * ldK + k] = acc;
      }
      // gz[i][f] = sum_k S[i][k] dX'[k][f] + readout scatter
      for (int idx = tid; idx < n * F; idx += blockDim.x) {
        const int il = idx / F, f = idx - il * F, i = rs + il;
        const float* s = bS + i * ldK;
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf(s[k], s_dxp[k * ldF + f], acc);
        const long long o = (long long)g * p.ldo + f;
        if (p.arg[o] == il) acc += p.dout[o];
        p.gz[(long long)(r0 + i) * F + f] = acc;
      }
    }
    __syncthreads();
    for (int i = wid; i < nt; i += nw) {                        // softmax backward: dT = s (dS - <dS, s>)
      float* d = bD + i * ldK;
      const float* s = bS + i * ldK;
      float pr = 0.f;
      for (int k = lane; k < K; k += 32) pr = fmaf(d[k], s[k], pr);
      pr = warp_sum(pr);
      for (int k = lane; k < ldK; k += 32) d[k] = k < K ? s[k] * (d[k] - pr) : 0.f;
    }
    __syncthreads();
    gemm_tn_acc(bD, ldK, bZa, ldFa, ldK >> 2, ldFa >> 2, 0, nt, s_dwp, ldFa, 1);
    if (p.dbp)
      for (int k = tid; k < K; k += blockDim.x) {
        float acc = 0.f;
        for (int i = 0; i < nt; ++i) acc += bD[i * ldK + k];
        s_dbp[k] += acc;
      }
    {
      float* gza = p.gza;
      gemm_nn(bD, ldK, s_wp, ldFa, nt, ldFa >> 2, ldK >> 2, [=](int m, int n, float a) {
        if (n < Fa) gza[(long long)(r0 + m) * Fa + n] = a;
      });
    }
  }
  __syncthreads();
  for (int idx = tid; idx < K * Fa; idx += blockDim.x) {
    const int k = idx / Fa, f = idx - k * Fa;
    const float v = s_dwp[k * ldFa + f];
    if (v != 0.f) atomicAdd(&p.dWp[idx], v);
  }
  if (p.dbp)
    for (int k = tid; k < K; k += blockDim.x)
      if (s_dbp[k] != 0.f) atomicAdd(&p.dbp[k], s_dbp[k]);
}

// ------------------------------------------------------------------------------------------------------------------
// max readout of a packed concat (pooled level)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
readout_kernel(const gp_pk_tiling tl, const gp_pk_concat z, const float* __restrict__ cnt_pad,
               const int32_t* __restrict__ nb, int N, float* __restrict__ out, int32_t* __restrict__ arg,
               long long ldo, int ooff) {
  extern __shared__ __align__(16) float sm[];
  float* s_bn = sm;
  concat_stats(z, cnt_pad, N, tl.B, s_bn);
  __syncthreads();
  const int F = z.F;
  const long long total = (long long)tl.B * F;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx / F), f = (int)(idx - (long long)g * F);
    int l = 0, off = 0;
    while (l < z.L - 1 && f >= off + z.slot[l].d) { off += z.slot[l].d; ++l; }
    const gp_pk_src& src = z.slot[l];
    const int c = f - off;
    const int rs = t_row0(tl, g), n = t_row0(tl, g + 1) - rs;
    const float* mean = s_bn + l * 2 * N;
    const float* istd = mean + N;
    float best = -INFINITY;
    int a = -1;
    for (int i = 0; i < n; ++i) {
      float v = src.y[(long long)(rs + i) * src.ld + c];
      if (src.sums) v = (fmaxf(v, 0.f) - mean[i]) * istd[i];
      if (v > best) { best = v; a = i; }
    }
    if (n < N && !(best >= 0.f)) { best = 0.f; a = -1; }
    out[(long long)g * ldo + ooff + f] = best;
    arg[(long long)g * ldo + ooff + f] = a;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// link-prediction loss on the real blocks (one CTA per graph at a time); P and dl/dP live in shared memory only
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float link_dot(const float* a, const float* b, int ldK) {
  float acc = 0.f;
  for (int k = 0; k < ldK; k += 4) {
    const float4 x = *reinterpret_cast<const float4*>(a + k);
    const float4 y = *reinterpret_cast<const float4*>(b + k);
    acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
  }
  return acc;
}

__global__ void __launch_bounds__(kThreads)
link_fwd_kernel(const float* __restrict__ S, const float* __restrict__ adj, const int32_t* __restrict__ nb, int B,
                int N, int K, double* __restrict__ sum) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float s_red[33];
  const int ldK = r4(K) + 4;                   // +4: rows land on different banks
  float* s_s = sm;
  const int tid = threadIdx.x;
  double tot = 0.0;
  for (int g = blockIdx.x; g < B; g += gridDim.x) {
    const int n = nb ? min(max(nb[g], 0), N) : N;
    __syncthreads();
    for (int idx = tid; idx < n * ldK; idx += blockDim.x) {
      const int i = idx / ldK, k = idx - i * ldK;
      s_s[idx] = k < K ? S[((long long)g * N + i) * K + k] : 0.f;
    }
    __syncthreads();
    float loc = 0.f;
    const float* ag = adj + (long long)g * N * N;
    for (int idx = tid; idx < n * n; idx += blockDim.x) {
      const int i = idx / n, j = idx - i * n;
      const float pr = fminf(link_dot(s_s + i * ldK, s_s + j * ldK, ldK), 1.f);
      const float a = ag[(long long)i * N + j];
      loc += -a * logf(pr + kEpsLink) - (1.f - a) * logf(1.f - pr + kEpsLink);
    }
    const float t = block_sum(loc, s_red);
    tot += (double)t;
  }
  if (tid == 0 && tot != 0.0) atomicAdd(sum, tot);
}

__global__ void __launch_bounds__(kThreads)
link_bwd_kernel(const float* __restrict__ S, const float* __restrict__ adj, const int32_t* __restrict__ nb, int B,
                int N, int K, float alpha, const float* __restrict__ alpha_dev, const float* __restrict__ alpha_dev2,
                float* __restrict__ dS) {
  extern __shared__ __align__(16) float sm[];
  const int ldK = r4(K) + 4;
  const int ldG = r4(N);
  float* s_s = sm;                 // [N x ldK]
  float* s_g = sm + r4(N) * ldK;   // [N x ldG]   G + G^T
  const int tid = threadIdx.x;
  float sc = alpha;
  if (alpha_dev) sc *= *alpha_dev;
  if (alpha_dev2) sc *= *alpha_dev2;
  for (int g = blockIdx.x; g < B; g += gridDim.x) {
    const int n = nb ? min(max(nb[g], 0), N) : N;
    const int n4 = r4(n);
    __syncthreads();
    for (int idx = tid; idx < n4 * ldK; idx += blockDim.x) {
      const int i = idx / ldK, k = idx - i * ldK;
      s_s[idx] = (i < n && k < K) ? S[((long long)g * N + i) * K + k] : 0.f;
    }
    __syncthreads();
    const float* ag = adj + (long long)g * N * N;
    for (int idx = tid; idx < n * n4; idx += blockDim.x) {
      const int i = idx / n4, j = idx - i * n4;
      float gsum = 0.f;
      if (j < n) {
        const float praw = link_dot(s_s + i * ldK, s_s + j * ldK, ldK);
        if (praw <= 1.f) {                      // clamp(max = 1) passes no gradient above 1
          const float aij = ag[(long long)i * N + j], aji = ag[(long long)j * N + i];
          const float u = 1.f / (praw + kEpsLink), v = 1.f / (1.f - praw + kEpsLink);
          gsum = (-aij * u + (1.f - aij) * v) + (-aji * u + (1.f - aji) * v);
        }
      }
      s_g[i * ldG + j] = gsum;
    }
    __syncthreads();
    float* o = dS + (long long)g * N * K;
    gemm_nn(s_g, ldG, s_s, ldK, n, (ldK - 4) >> 2, n4 >> 2, [=](int m, int c, float a) {
      if (c < K) o[(long long)m * K + c] = sc * a;
    });
    for (int idx = tid; idx < (N - n) * K; idx += blockDim.x) o[(long long)n * K + idx] = 0.f;
  }
}

__global__ void link_finalize_kernel(const double* sum, double scale, const float* scale_dev, const float* base,
                                     float* total, float* link) {
  double l = *sum * scale;
  if (scale_dev) l *= (double)*scale_dev;
  if (link) *link = (float)l;
  if (total) *total = (float)((base ? (double)*base : 0.0) + l);
}

static int grid_for(int smem_bytes, int want) {
  // persistent CTAs: as many as fit per SM (228 KB of shared memory per SM, 1 KB reserved per CTA), at most `want`
  int per_sm = (int)(228 * 1024 / (smem_bytes + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  const int g = kNumSMs * per_sm;
  return g < want ? g : (want < 1 ? 1 : want);
}

// dynamic shared memory opt-in, once per device and kernel (not a stream operation: safe under graph capture too)
#define GP_PK_SMEM(kernel, bytes)                                                                              \
  do {                                                                                                         \
    if ((bytes) > 226 * 1024)                                                                                  \
      return gp::fail(GP_ERR_UNSUPPORTED, "packed schedule: %zu bytes of shared memory", (size_t)(bytes));     \
    GP_CONFIG_ONCE(GP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024))); \
  } while (0)

static int tiles_upper(const gp_pk_tiling& t) {
  if (t.tile_g0 == nullptr) return (t.B + t.gpt - 1) / t.gpt;
  return t.gpt;            // ragged tiling: the caller passes an upper bound of the tile count in gpt
}

}  // namespace pk
}  // namespace gp

using namespace gp;
using namespace gp::pk;

extern "C" int gp_pk_prepare(const int32_t* nb, int B, int N, int w_layer, int w_pool, int32_t* rowptr, float* cnt_pad,
                             int32_t* tiles_layer, int32_t* tiles_pool, int32_t* meta, gp_stream_t stream) {
  GP_REQUIRE(rowptr && cnt_pad && tiles_layer && tiles_pool && meta, "pk_prepare: null pointer");
  GP_REQUIRE(B > 0 && N > 0 && N <= kMaxN && w_layer > 0 && w_pool > 0, "pk_prepare: B=%d N=%d (N <= %d)", B, N, kMaxN);
  prepare_kernel<<<1, 1024, 0, S(stream)>>>(nb, B, N, w_layer, w_pool, rowptr, cnt_pad, tiles_layer, tiles_pool, meta);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_pk_build_lists(const float* adj, const int32_t* nb, const int32_t* rowptr, int B, int N,
                                 int32_t* info_out, int32_t* ent_out, int32_t* info_in, int32_t* ent_in,
                                 int32_t* cursor, long long capacity, gp_stream_t stream) {
  GP_REQUIRE(adj && rowptr && info_out && ent_out && info_in && ent_in && cursor, "pk_build_lists: null pointer");
  GP_REQUIRE(B > 0 && N > 0 && N <= kMaxN && capacity > 0 && capacity < (1ll << 31), "pk_build_lists: bad sizes");
  const size_t smem = (size_t)(N * (N + 1) + 2 * (N + 1)) * sizeof(float);
  GP_PK_SMEM(build_lists_kernel, smem);
  const int grid = grid_for((int)smem, B);
  build_lists_kernel<<<grid, 128, smem, S(stream)>>>(adj, nb, rowptr, B, N, reinterpret_cast<int2*>(info_out),
                                                     reinterpret_cast<int2*>(ent_out), reinterpret_cast<int2*>(info_in),
                                                     reinterpret_cast<int2*>(ent_in), cursor, capacity);
  GP_LAUNCHED();
  return GP_OK;
}

static int check_tiling(const gp_pk_tiling& t, int N) {
  GP_REQUIRE(t.B > 0 && t.max_rows > 0 && t.gpt > 0, "pk: bad tiling");
  GP_REQUIRE((t.rowptr != nullptr) == (t.tile_g0 != nullptr) && (t.rowptr != nullptr) == (t.ntiles != nullptr),
             "pk: ragged tiling needs rowptr, tile_g0 and ntiles");
  GP_REQUIRE(t.rowptr != nullptr || (t.nfix > 0 && t.nfix <= N && t.gpt * t.nfix <= t.max_rows), "pk: uniform tiling");
  GP_REQUIRE(N > 0 && N <= kMaxN, "pk: N = %d (<= %d)", N, kMaxN);
  return GP_OK;
}
static int check_adj(const gp_pk_adj& a, const gp_pk_tiling& t) {
  GP_REQUIRE((a.info && a.entries) || (a.dense && t.rowptr == nullptr), "pk: adjacency (lists, or dense with a uniform tiling)");
  return GP_OK;
}

extern "C" int gp_pk_layer_fwd(const gp_pk_layer_fwd_args* a, gp_stream_t stream) {
  GP_REQUIRE(a != nullptr, "pk_layer_fwd: null args");
  GP_TRY(check_tiling(a->tl, a->N));
  GP_TRY(check_adj(a->adj, a->tl));
  GP_REQUIRE(a->ns == 1 || a->ns == 2, "pk_layer_fwd: ns");
  for (int s = 0; s < a->ns; ++s) {
    const gp_pk_stack_fwd& st = a->s[s];
    GP_REQUIRE(st.in.y && st.W && st.y && st.rnorm && st.in.d > 0 && st.dout > 0 && st.in.d <= 512 && st.dout <= 512,
               "pk_layer_fwd: stack %d", s);
  }
  const FwdDims dm = fwd_dims(*a);
  Count c;
  fwd_carve(*a, dm, c);
  const size_t smem = c.n * sizeof(float);
  GP_PK_SMEM(layer_fwd_kernel, smem);
  layer_fwd_kernel<<<grid_for((int)smem, tiles_upper(a->tl)), kThreads, smem, S(stream)>>>(*a);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_pk_layer_bwd(const gp_pk_layer_bwd_args* a, gp_stream_t stream) {
  GP_REQUIRE(a != nullptr, "pk_layer_bwd: null args");
  GP_TRY(check_tiling(a->tl, a->N));
  GP_TRY(check_adj(a->adj, a->tl));
  GP_REQUIRE(a->ns == 1 || a->ns == 2, "pk_layer_bwd: ns");
  for (int s = 0; s < a->ns; ++s) {
    const gp_pk_stack_bwd& st = a->s[s];
    GP_REQUIRE(st.in.y && st.out.y && st.W && st.rnorm && st.dW && st.in.d > 0 && st.dout > 0, "pk_layer_bwd: stack %d", s);
    GP_REQUIRE(st.out.d == st.dout, "pk_layer_bwd: out.d != dout");
    GP_REQUIRE((st.out.sums == nullptr) == (st.msums == nullptr), "pk_layer_bwd: msums go with a BatchNorm'd output");
    GP_REQUIRE(st.gl.dense || st.gl.dout, "pk_layer_bwd: no upstream gradient");
    if (st.need_dx) {
      GP_REQUIRE(st.gl_prev != nullptr, "pk_layer_bwd: gl_prev");
      GP_TRY(check_adj(a->adj_in, a->tl));
    }
    GP_REQUIRE(st.dadj == nullptr || a->tl.rowptr == nullptr, "pk_layer_bwd: dadj needs the dense level");
  }
  const BwdDims dm = bwd_dims(*a);
  Count c;
  bwd_carve(*a, dm, c);
  const size_t smem = c.n * sizeof(float);
  GP_PK_SMEM(layer_bwd_kernel, smem);
  layer_bwd_kernel<<<grid_for((int)smem, tiles_upper(a->tl)), kThreads, smem, S(stream)>>>(*a);
  GP_LAUNCHED();
  return GP_OK;
}

static int check_pool(const gp_pk_pool_args* a, bool bwd) {
  GP_REQUIRE(a != nullptr, "pk_pool: null args");
  GP_TRY(check_tiling(a->tl, a->N));
  GP_TRY(check_adj(a->adj, a->tl));
  GP_TRY(check_adj(a->adj_in, a->tl));
  GP_REQUIRE(a->K > 0 && a->K <= kMaxN && a->z.L > 0 && a->z.L <= GP_PK_MAX_LAYERS && a->za.L > 0 &&
             a->za.L <= GP_PK_MAX_LAYERS, "pk_pool: K / layer counts");
  int f = 0, fa = 0;
  for (int l = 0; l < a->z.L; ++l) { GP_REQUIRE(a->z.slot[l].y, "pk_pool: z slot"); f += a->z.slot[l].d; }
  for (int l = 0; l < a->za.L; ++l) { GP_REQUIRE(a->za.slot[l].y, "pk_pool: za slot"); fa += a->za.slot[l].d; }
  GP_REQUIRE(f == a->z.F && fa == a->za.F, "pk_pool: concat widths");
  GP_REQUIRE(a->Wp && a->S && a->arg, "pk_pool: null pointer");
  if (!bwd) GP_REQUIRE(a->xp && a->ap && a->out, "pk_pool_fwd: null output");
  else GP_REQUIRE(a->dxp && a->dap && a->dout && a->gz && a->gza && a->dWp, "pk_pool_bwd: null pointer");
  return GP_OK;
}

extern "C" int gp_pk_pool_fwd(const gp_pk_pool_args* a, gp_stream_t stream) {
  GP_TRY(check_pool(a, false));
  const PoolDims dm = pool_dims(*a);
  Count c;
  pool_carve(*a, dm, false, c);
  const size_t smem = c.n * sizeof(float);
  GP_PK_SMEM(pool_fwd_kernel, smem);
  pool_fwd_kernel<<<grid_for((int)smem, tiles_upper(a->tl)), kThreads, smem, S(stream)>>>(*a);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_pk_pool_bwd(const gp_pk_pool_args* a, gp_stream_t stream) {
  GP_TRY(check_pool(a, true));
  const PoolDims dm = pool_dims(*a);
  Count c;
  pool_carve(*a, dm, true, c);
  const size_t smem = c.n * sizeof(float);
  GP_PK_SMEM(pool_bwd_kernel, smem);
  pool_bwd_kernel<<<grid_for((int)smem, tiles_upper(a->tl)), kThreads, smem, S(stream)>>>(*a);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_pk_readout(const gp_pk_tiling* tl, const gp_pk_concat* z, const float* cnt_pad, const int32_t* nb,
                             int N, float* out, int32_t* arg, long long ldo, int ooff, gp_stream_t stream) {
  GP_REQUIRE(tl && z && out && arg, "pk_readout: null pointer");
  GP_REQUIRE(N > 0 && N <= kMaxN && z->L > 0 && z->L <= GP_PK_MAX_LAYERS && tl->B > 0, "pk_readout: sizes");
  GP_REQUIRE(tl->rowptr != nullptr || tl->nfix > 0, "pk_readout: tiling");
  const size_t smem = (size_t)z->L * 2 * N * sizeof(float);
  const long long total = (long long)tl->B * z->F;
  const int grid = (int)((total + kThreads - 1) / kThreads < 4 * kNumSMs ? (total + kThreads - 1) / kThreads : 4 * kNumSMs);
  readout_kernel<<<grid, kThreads, smem, S(stream)>>>(*tl, *z, cnt_pad, nb, N, out, arg, ldo, ooff);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_pk_link_fwd(const float* Sm, const float* adj, const int32_t* nb, int B, int N, int K, double* sum,
                              gp_stream_t stream) {
  GP_REQUIRE(Sm && adj && sum && B > 0 && N > 0 && N <= kMaxN && K > 0 && K <= kMaxN, "pk_link_fwd: bad args");
  const size_t smem = (size_t)N * (r4(K) + 4) * sizeof(float);
  GP_PK_SMEM(link_fwd_kernel, smem);
  link_fwd_kernel<<<grid_for((int)smem, B), kThreads, smem, S(stream)>>>(Sm, adj, nb, B, N, K, sum);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_pk_link_bwd(const float* Sm, const float* adj, const int32_t* nb, int B, int N, int K, float alpha,
                              const float* alpha_dev, const float* alpha_dev2, float* dS, gp_stream_t stream) {
  GP_REQUIRE(Sm && adj && dS && B > 0 && N > 0 && N <= kMaxN && K > 0 && K <= kMaxN, "pk_link_bwd: bad args");
  const size_t smem = (size_t)(r4(N) * (r4(K) + 4) + r4(N) * r4(N)) * sizeof(float);
  GP_PK_SMEM(link_bwd_kernel, smem);
  link_bwd_kernel<<<grid_for((int)smem, B), kThreads, smem, S(stream)>>>(Sm, adj, nb, B, N, K, alpha, alpha_dev,
                                                                        alpha_dev2, dS);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_pk_link_finalize(const double* sum, double scale, const float* scale_dev, const float* base,
                                   float* total, float* link, gp_stream_t stream) {
  GP_REQUIRE(sum && (total || link), "pk_link_finalize: bad args");
  link_finalize_kernel<<<1, 1, 0, S(stream)>>>(sum, scale, scale_dev, base, total, link);
  GP_LAUNCHED();
  return GP_OK;
}
