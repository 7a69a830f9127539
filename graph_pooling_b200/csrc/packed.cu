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
// runs ("sub-tiles"): groups of whole graphs with at most max_rows packed rows; one CTA processes a run at a time
// ------------------------------------------------------------------------------------------------------------------
struct Sub { int r0, nt, g0, g1; };
__device__ __forceinline__ int t_row0(const gp_pk_tiling& t, int g) { return t.rowptr ? t.rowptr[g] : g * t.nfix; }
__device__ __forceinline__ int sub_count(const gp_pk_tiling& t) {
  return t.subs ? *t.nsub : (t.B + t.gpt - 1) / t.gpt;
}
__device__ __forceinline__ Sub sub_get(const gp_pk_tiling& t, int s) {
  Sub r;
  if (t.subs) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(t.subs) + s);
    r.r0 = v.x; r.nt = v.y; r.g0 = v.z; r.g1 = v.w;
  } else {
    r.g0 = s * t.gpt;
    r.g1 = min(t.B, r.g0 + t.gpt);
    r.r0 = r.g0 * t.nfix;
    r.nt = (r.g1 - r.g0) * t.nfix;
  }
  return r;
}
// per-row maps of a run: s_gs[i] = local row of the first row of i's graph, s_gid[i] = its graph
__device__ __forceinline__ void load_meta(const gp_pk_tiling& t, const Sub& sb, int* s_gs, int* s_gid) {
  for (int i = threadIdx.x; i < sb.nt; i += blockDim.x) {
    int ni, gid;
    if (t.rowmeta) {
      const int2 m = __ldg(reinterpret_cast<const int2*>(t.rowmeta) + sb.r0 + i);
      ni = m.x;
      gid = m.y;
    } else {
      const int q = i / t.nfix;
      ni = i - q * t.nfix;
      gid = sb.g0 + q;
    }
    s_gs[i] = i - ni;
    s_gid[i] = gid;
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
// One warp per row, lane = column (no index arithmetic per element), four rows in flight per warp.
__device__ void load_rows(const gp_pk_src& src, int r0, int nt, const int* s_gs, const float* s_mean,
                          const float* s_istd, float* dst, int ldd, int coff, int dz) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int d = src.d;
  const bool bn = src.sums != nullptr;
  const float* base = src.y + (long long)r0 * src.ld;
  const int ld = (int)src.ld;
  for (int i0 = wid; i0 < nt; i0 += 4 * nw) {
    for (int c = lane; c < dz; c += 32) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * nw;
        v[j] = (i < nt && c < d) ? __ldg(base + i * ld + c) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * nw;
        if (i < nt) {
          float x = v[j];
          if (bn && c < d) {
            const int ni = i - s_gs[i];
            x = (fmaxf(x, 0.f) - s_mean[ni]) * s_istd[ni];
          }
          dst[i * ldd + coff + c] = x;
        }
      }
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
// prepare: scan of the node counts, pad counts per node index, the run table
// ------------------------------------------------------------------------------------------------------------------
// first graph whose first row is >= target
__device__ __forceinline__ int first_graph_at(const int32_t* rowptr, int B, int target) {
  int lo = 0, hi = B;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (rowptr[mid] < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// greedy split of the graphs [G0, G1) of one window into runs of at most `cap` rows; returns the number of runs and,
// with out != nullptr, writes them
__device__ __forceinline__ int split_window(const int32_t* rowptr, int G0, int G1, int cap, int4* out) {
  int cnt = 0;
  for (int g = G0; g < G1;) {
    const int r = rowptr[g];
    int ge = g + 1;
    while (ge < G1 && rowptr[ge + 1] - r <= cap) ++ge;
    if (out) out[cnt] = make_int4(r, rowptr[ge] - r, g, ge);
    ++cnt;
    g = ge;
  }
  return cnt;
}

__global__ void __launch_bounds__(1024)
prepare_kernel(const int32_t* __restrict__ nb, int B, int N, int w, int cap, int32_t* __restrict__ rowptr,
               float* __restrict__ cnt_pad, int32_t* __restrict__ subs, int32_t* __restrict__ meta) {
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
  if (tid == 0) {                               // cnt_pad[n] = #graphs with n_b <= n
    int c = 0;
    for (int n = 0; n < N; ++n) {
      c += s_hist[n];
      cnt_pad[n] = (float)c;
    }
  }
  __threadfence_block();
  __syncthreads();
  // windows of w packed rows -> runs; two passes (count, scan, write) so that the table is dense
  const int nwin = (R + w - 1) / w;
  const int wper = (nwin + nth - 1) / nth;
  const int w0 = min(nwin, tid * wper), w1 = min(nwin, w0 + wper);
  int mine = 0;
  for (int t = w0; t < w1; ++t) {
    const int G0 = first_graph_at(rowptr, B, t * w);
    const int G1 = t + 1 < nwin ? first_graph_at(rowptr, B, (t + 1) * w) : B;
    mine += split_window(rowptr, G0, G1, cap, nullptr);
  }
  __syncthreads();
  s_part[tid] = mine;
  __syncthreads();
  for (int o = 1; o < nth; o <<= 1) {
    const int v = tid >= o ? s_part[tid - o] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  int at = tid ? s_part[tid - 1] : 0;
  for (int t = w0; t < w1; ++t) {
    const int G0 = first_graph_at(rowptr, B, t * w);
    const int G1 = t + 1 < nwin ? first_graph_at(rowptr, B, (t + 1) * w) : B;
    at += split_window(rowptr, G0, G1, cap, reinterpret_cast<int4*>(subs) + at);
  }
  if (tid == 0) { meta[0] = R; meta[1] = s_part[nth - 1]; meta[2] = 0; meta[3] = 0; }
}

// ------------------------------------------------------------------------------------------------------------------
// neighbour lists of the dense level-0 adjacency (one CTA per graph at a time): ELL part (first kEll entries of every
// row at a fixed position, zero padded: loadable without knowing the degree) + overflow; rowmeta; packed inputs
// ------------------------------------------------------------------------------------------------------------------
constexpr int kEll = GP_PK_ELL;

__global__ void __launch_bounds__(128)
build_lists_kernel(const float* __restrict__ adj, const int32_t* __restrict__ nb, const int32_t* __restrict__ rowptr,
                   int B, int N, int2* __restrict__ info_out, int2* __restrict__ ell_out, int2* __restrict__ ovf_out,
                   int2* __restrict__ info_in, int2* __restrict__ ell_in, int2* __restrict__ ovf_in,
                   int32_t* __restrict__ cursors, long long capacity, int2* __restrict__ rowmeta,
                   const float* __restrict__ x, int D, float* __restrict__ xpack, long long ldxp,
                   const float* __restrict__ ax, int Da, float* __restrict__ axpack, long long ldaxp) {
  extern __shared__ __align__(16) float sm[];
  float* s_a = sm;                                           // [n][n+1]
  int* s_do = reinterpret_cast<int*>(sm + N * (N + 1));      // out degree, then overflow offset [N+1]
  int* s_di = s_do + N + 1;                                  // in degree, then overflow offset
  __shared__ int s_base[2];
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
    // packed copies of the inputs and the row map
    for (int i = tid; i < n; i += blockDim.x) rowmeta[r0 + i] = make_int2(i, g);
    if (xpack)
      for (int idx = tid; idx < n * (int)ldxp; idx += blockDim.x) {
        const int i = idx / (int)ldxp, c = idx - i * (int)ldxp;
        xpack[(long long)(r0 + i) * ldxp + c] = c < D ? x[((long long)g * N + i) * D + c] : 0.f;
      }
    if (axpack)
      for (int idx = tid; idx < n * (int)ldaxp; idx += blockDim.x) {
        const int i = idx / (int)ldaxp, c = idx - i * (int)ldaxp;
        axpack[(long long)(r0 + i) * ldaxp + c] = c < Da ? ax[((long long)g * N + i) * Da + c] : 0.f;
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
    if (tid == 0) {                                          // overflow offsets (entries beyond the ELL part)
      int ro = 0, ri = 0;
      for (int i = 0; i < n; ++i) {
        ro += max(s_do[i] - kEll, 0);
        ri += max(s_di[i] - kEll, 0);
      }
      s_base[0] = ro > 0 ? atomicAdd(cursors, ro) : 0;
      s_base[1] = ri > 0 ? atomicAdd(cursors + 1, ri) : 0;
      ro = s_base[0];
      ri = s_base[1];
      for (int i = 0; i < n; ++i) {
        const int a = s_do[i], b = s_di[i];
        // degree in the low half, overflow position kept in the info record
        s_do[i] = ro;
        s_di[i] = ri;
        ro += max(a - kEll, 0);
        ri += max(b - kEll, 0);
        info_out[r0 + i] = make_int2(s_do[i], (long long)ro <= capacity ? a : min(a, kEll));
        info_in[r0 + i] = make_int2(s_di[i], (long long)ri <= capacity ? b : min(b, kEll));
      }
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
      int po = 0, pi = 0;
      const int degs_o = info_out[r0 + i].y, degs_i = info_in[r0 + i].y;
      int2* eo = ell_out + (long long)(r0 + i) * kEll;
      int2* ei = ell_in + (long long)(r0 + i) * kEll;
      for (int j = 0; j < n; ++j) {
        const float vo = s_a[i * ld + j];
        if (vo != 0.f && po < degs_o) {
          if (po < kEll) eo[po] = make_int2(j, __float_as_int(vo));
          else ovf_out[s_do[i] + po - kEll] = make_int2(j, __float_as_int(vo));
          ++po;
        }
        const float vi = s_a[j * ld + i];
        if (vi != 0.f && pi < degs_i) {
          if (pi < kEll) ei[pi] = make_int2(j, __float_as_int(vi));
          else ovf_in[s_di[i] + pi - kEll] = make_int2(j, __float_as_int(vi));
          ++pi;
        }
      }
      for (; po < kEll; ++po) eo[po] = make_int2(0, 0);
      for (; pi < kEll; ++pi) ei[pi] = make_int2(0, 0);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// shared pieces of the layer / pooling kernels
// ------------------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int pow2_ge(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Neighbour lists (or the dense blocks of the pooled level) of a run -> shared memory.  Every load is independent of
// every other (the ELL part sits at a fixed position per row), so the whole run's lists are in flight at once: one
// global round trip per run instead of two dependent ones per row.   s_info [mr] int2, s_ell [mr * kEll] int2.
__device__ __forceinline__ bool dense_fits(const gp_pk_tiling& tl, const Sub& sb) {
  return (sb.g1 - sb.g0) * tl.nfix * tl.nfix <= 2 * tl.max_rows * kEll;
}
__device__ __forceinline__ void stage_lists(const gp_pk_adj& a, const gp_pk_tiling& tl, const Sub& sb, int2* s_info,
                                            int2* s_ell) {
  if (a.info) {
    const int2* gi = reinterpret_cast<const int2*>(a.info) + sb.r0;
    for (int i = threadIdx.x; i < sb.nt; i += blockDim.x) s_info[i] = __ldg(gi + i);
    const int4* ge = reinterpret_cast<const int4*>(a.ell) + (size_t)sb.r0 * (kEll / 2);
    int4* se = reinterpret_cast<int4*>(s_ell);
    const int total = sb.nt * (kEll / 2);
    for (int base = threadIdx.x; base < total; base += 4 * blockDim.x) {
      int4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (base + j * blockDim.x < total) v[j] = __ldg(ge + base + j * blockDim.x);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (base + j * blockDim.x < total) se[base + j * blockDim.x] = v[j];
    }
  } else if (dense_fits(tl, sb)) {
    const int per = tl.nfix * tl.nfix;
    float* d = reinterpret_cast<float*>(s_ell);
    const float* src = a.dense + (long long)sb.g0 * per;
    for (int idx = threadIdx.x; idx < (sb.g1 - sb.g0) * per; idx += blockDim.x) d[idx] = __ldg(src + idx);
  }
}

// u = sum_e a(i, e) * rows[gs + col_e][lane]        (one warp per row, lane = column, ld <= 32)
__device__ __forceinline__ float gather_lane(const gp_pk_adj& a, const gp_pk_tiling& tl, const Sub& sb, int i, int gs,
                                             int gid, int ni, const int2* s_info, const int2* s_ell,
                                             const float* rows, int ld, int lane) {
  float u = 0.f;
  const float* sp = rows + (size_t)gs * ld + lane;
  if (a.info) {
    const int2 inf = s_info[i];
    if (lane < ld) {
      const int m = min(inf.y, kEll);
      for (int e = 0; e < m; ++e) {
        const int2 x = s_ell[i * kEll + e];
        u = fmaf(__int_as_float(x.y), sp[(size_t)x.x * ld], u);
      }
      const int2* en = reinterpret_cast<const int2*>(a.entries) + inf.x;
      for (int e = kEll; e < inf.y; ++e) {
        const int2 x = __ldg(en + e - kEll);
        u = fmaf(__int_as_float(x.y), sp[(size_t)x.x * ld], u);
      }
    }
  } else {
    const int nf = tl.nfix, per = nf * nf;
    const float* dn = dense_fits(tl, sb) ? reinterpret_cast<const float*>(s_ell) + (gid - sb.g0) * per
                                         : a.dense + (long long)gid * per;
    if (lane < ld)
      for (int e = 0; e < nf; ++e) {
        const float v = a.transposed ? dn[e * nf + ni] : dn[ni * nf + e];
        u = fmaf(v, sp[(size_t)e * ld], u);
      }
  }
  return u;
}

// sum_e a(i, e) * src[gs + col_e][4q .. 4q+3]
__device__ __forceinline__ float4 gather_row4(const gp_pk_adj& a, const gp_pk_tiling& tl, const Sub& sb, int i, int gs,
                                              int gid, const int2* s_info, const int2* s_ell, const float* src, int ld,
                                              int q) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int ni = i - gs;
  const float* sp = src + (size_t)gs * ld + 4 * q;
  if (a.info) {
    const int2 inf = s_info[i];
    const int m = min(inf.y, kEll);
    for (int e = 0; e < m; ++e) {
      const int2 x = s_ell[i * kEll + e];
      const float v = __int_as_float(x.y);
      const float4 t = *reinterpret_cast<const float4*>(sp + (size_t)x.x * ld);
      acc.x = fmaf(v, t.x, acc.x); acc.y = fmaf(v, t.y, acc.y); acc.z = fmaf(v, t.z, acc.z); acc.w = fmaf(v, t.w, acc.w);
    }
    const int2* en = reinterpret_cast<const int2*>(a.entries) + inf.x;
    for (int e = kEll; e < inf.y; ++e) {
      const int2 x = __ldg(en + e - kEll);
      const float v = __int_as_float(x.y);
      const float4 t = *reinterpret_cast<const float4*>(sp + (size_t)x.x * ld);
      acc.x = fmaf(v, t.x, acc.x); acc.y = fmaf(v, t.y, acc.y); acc.z = fmaf(v, t.z, acc.z); acc.w = fmaf(v, t.w, acc.w);
    }
  } else {
    const int nf = tl.nfix, per = nf * nf;
    const float* dn = dense_fits(tl, sb) ? reinterpret_cast<const float*>(s_ell) + (gid - sb.g0) * per
                                         : a.dense + (long long)gid * per;
    for (int e = 0; e < nf; ++e) {
      const float v = a.transposed ? dn[e * nf + ni] : dn[ni * nf + e];
      const float4 t = *reinterpret_cast<const float4*>(sp + (size_t)e * ld);
      acc.x = fmaf(v, t.x, acc.x); acc.y = fmaf(v, t.y, acc.y); acc.z = fmaf(v, t.z, acc.z); acc.w = fmaf(v, t.w, acc.w);
    }
  }
  return acc;
}
// dst = A src over a run, one thread per (row, 4-column chunk)
__device__ void gather4(const gp_pk_adj& a, const gp_pk_tiling& tl, const Sub& sb, const int* s_gs, const int* s_gid,
                        const int2* s_info, const int2* s_ell, const float* src, int ld, float* dst) {
  const int C4 = ld >> 2;
  for (int item = threadIdx.x; item < sb.nt * C4; item += blockDim.x) {
    const int i = item / C4, q = item - i * C4;
    const float4 acc = gather_row4(a, tl, sb, i, s_gs[i], s_gid[i], s_info, s_ell, src, ld, q);
    *reinterpret_cast<float4*>(dst + (size_t)i * ld + 4 * q) = acc;
  }
}

// rows of the run -> shared memory, one warp per row (ld <= 32), four rows in flight per warp
__device__ __forceinline__ void load_rows_warp(const gp_pk_src& src, const Sub& sb, const int* s_gs,
                                               const float* s_mean, const float* s_istd, float* dst, int ld) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const bool bn = src.sums != nullptr;
  if (lane >= ld) return;
#pragma unroll 4
  for (int i = wid; i < sb.nt; i += nw) {
    float v = 0.f;
    if (lane < src.d) {
      v = __ldg(src.y + (long long)(sb.r0 + i) * src.ld + lane);
      if (bn) {
        const int ni = i - s_gs[i];
        v = (fmaxf(v, 0.f) - s_mean[ni]) * s_istd[ni];
      }
    }
    dst[i * ld + lane] = v;
  }
}

// Deterministic per-node-index accumulation of per-row statistics: s_rs[i], s_rs[mr + i] hold the two values of local
// row i; thread n adds the rows with node index n graph by graph (fixed order), so a forward pass is bit-reproducible
// (shared-memory atomics would add them in scheduling order).
__device__ __forceinline__ void reduce_row_stats(const gp_pk_tiling& tl, const Sub& sb, int N, int mr,
                                                 const float* s_rs, float* s_st) {
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float a1 = 0.f, a2 = 0.f;
    for (int g = sb.g0; g < sb.g1; ++g) {
      const int gs = t_row0(tl, g) - sb.r0, ng = t_row0(tl, g + 1) - sb.r0 - gs;
      if (n < ng) {
        a1 += s_rs[gs + n];
        a2 += s_rs[mr + gs + n];
      }
    }
    s_st[n] += a1;
    s_st[N + n] += a2;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// GCN layer forward: blockIdx.y = stack (embedding / assignment GCN in lock-step), persistent over the window tiles
// ------------------------------------------------------------------------------------------------------------------
static size_t fwd_smem(const gp_pk_layer_fwd_args& p) {
  size_t best = 0;
  for (int s = 0; s < p.ns; ++s) {
    Count c;
    const int ldi = r4(p.s[s].in.d), ldo = r4(p.s[s].dout), mr = p.tl.max_rows;
    c.take(mr * ldi); c.take(mr * ldi); c.take(ldi * ldo); c.take(ldo); c.take(2 * p.N); c.take(2 * p.N);
    c.take(mr); c.take(mr); c.take(2 * mr); c.take(2 * mr); c.take(2 * mr * kEll);
    best = best > c.n ? best : c.n;
  }
  return best * sizeof(float);
}

__global__ void __launch_bounds__(kThreads, 3)
layer_fwd_kernel(const gp_pk_layer_fwd_args p) {
  extern __shared__ __align__(16) float sm[];
  const gp_pk_stack_fwd& st = p.s[blockIdx.y];
  const int N = p.N, din = st.in.d, dout = st.dout, ldi = r4(din), ldo = r4(dout), mr = p.tl.max_rows;
  Carve cv{sm};
  float* bH = cv.take(mr * ldi);
  float* bU = cv.take(mr * ldi);
  float* s_w = cv.take(ldi * ldo);
  float* s_b = cv.take(ldo);
  float* s_bn = cv.take(2 * N);
  float* s_st = cv.take(2 * N);
  int* s_gs = reinterpret_cast<int*>(cv.take(mr));
  int* s_gid = reinterpret_cast<int*>(cv.take(mr));
  float* s_rs = cv.take(2 * mr);
  int2* s_info = reinterpret_cast<int2*>(cv.take(2 * mr));
  int2* s_ell = reinterpret_cast<int2*>(cv.take(2 * mr * kEll));
  const int tid = threadIdx.x;
  for (int idx = tid; idx < ldi * ldo; idx += blockDim.x) {
    const int k = idx / ldo, n = idx - k * ldo;
    s_w[idx] = (k < din && n < dout) ? st.W[(long long)k * dout + n] : 0.f;
  }
  for (int n = tid; n < ldo; n += blockDim.x) s_b[n] = (st.b && n < dout) ? st.b[n] : 0.f;
  for (int n = tid; n < 2 * N; n += blockDim.x) s_st[n] = 0.f;
  bn_stats(st.in, p.cnt_pad, N, p.tl.B, s_bn, s_bn + N);
  const int N4 = ldo >> 2, N4p = pow2_ge(N4), K4 = ldi >> 2;
  const bool stats = st.sums_out != nullptr;
  const int nsub = sub_count(p.tl);
  for (int run = blockIdx.x; run < nsub; run += gridDim.x) {
    {
      const Sub sb = sub_get(p.tl, run);
      const int r0 = sb.r0, nt = sb.nt, g0 = sb.g0, g1 = sb.g1;
      (void)g0; (void)g1;
      if (nt > 0) {
        __syncthreads();
        load_meta(p.tl, sb, s_gs, s_gid);
        stage_lists(p.adj, p.tl, sb, s_info, s_ell);
        __syncthreads();
        load_rows(st.in, r0, nt, s_gs, s_bn, s_bn + N, bH, ldi, 0, ldi);
        __syncthreads();
        gather4(p.adj, p.tl, sb, s_gs, s_gid, s_info, s_ell, bH, ldi, bU);
        __syncthreads();
        // V = U W + b, Y = V / max(||V||, eps), ReLU statistics: a row lives in the N4p lanes of one aligned group
        const int total = ((nt + 3) >> 2) * N4p;
        for (int base = 0; base < total; base += blockDim.x) {
          const bool act = base + tid < total;
          const int item = act ? base + tid : total - 1;
          const int i = item / N4p, j = item - i * N4p;
          const bool cols = j < N4;
          const int mrem = nt - 4 * i;
          float acc[4][4];
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
          if (cols) {
            const float* a0 = bU + (size_t)(4 * i) * ldi;
            const float* a1 = a0 + (mrem > 1 ? ldi : 0);
            const float* a2 = a0 + (mrem > 2 ? 2 * ldi : 0);
            const float* a3 = a0 + (mrem > 3 ? 3 * ldi : 0);
            const float* b = s_w + 4 * j;
            for (int k4 = 0; k4 < K4; ++k4) {
              const float4 x0 = *reinterpret_cast<const float4*>(a0 + 4 * k4);
              const float4 x1 = *reinterpret_cast<const float4*>(a1 + 4 * k4);
              const float4 x2 = *reinterpret_cast<const float4*>(a2 + 4 * k4);
              const float4 x3 = *reinterpret_cast<const float4*>(a3 + 4 * k4);
              const float xs[4][4] = {{x0.x, x0.y, x0.z, x0.w}, {x1.x, x1.y, x1.z, x1.w}, {x2.x, x2.y, x2.z, x2.w},
                                      {x3.x, x3.y, x3.z, x3.w}};
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const float4 wv = *reinterpret_cast<const float4*>(b + (size_t)(4 * k4 + kk) * ldo);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                  acc[r][0] = fmaf(xs[r][kk], wv.x, acc[r][0]);
                  acc[r][1] = fmaf(xs[r][kk], wv.y, acc[r][1]);
                  acc[r][2] = fmaf(xs[r][kk], wv.z, acc[r][2]);
                  acc[r][3] = fmaf(xs[r][kk], wv.w, acc[r][3]);
                }
              }
            }
            const float4 bv = *reinterpret_cast<const float4*>(s_b + 4 * j);
#pragma unroll
            for (int r = 0; r < 4; ++r) { acc[r][0] += bv.x; acc[r][1] += bv.y; acc[r][2] += bv.z; acc[r][3] += bv.w; }
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            float ss = acc[r][0] * acc[r][0];
            ss = fmaf(acc[r][1], acc[r][1], ss);
            ss = fmaf(acc[r][2], acc[r][2], ss);
            ss = fmaf(acc[r][3], acc[r][3], ss);
            for (int o = N4p >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            const float rn = fmaxf(sqrtf(ss), kEpsNorm);
            float a1 = 0.f, a2 = 0.f;
            const bool rowok = act && r < mrem;
            float* yo = st.y + (long long)(r0 + 4 * i + r) * dout + 4 * j;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float y = acc[r][c] / rn;
              if (rowok && cols && 4 * j + c < dout) yo[c] = y;
              const float q = fmaxf(y, 0.f);
              a1 += q;
              a2 = fmaf(q, q, a2);
            }
            if (stats) {
              for (int o = N4p >> 1; o > 0; o >>= 1) {
                a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                a2 += __shfl_xor_sync(0xffffffffu, a2, o);
              }
            }
            if (rowok && j == 0) {
              st.rnorm[r0 + 4 * i + r] = rn;
              s_rs[4 * i + r] = a1;
              s_rs[mr + 4 * i + r] = a2;
            }
          }
        }
        if (stats) {
          __syncthreads();
          reduce_row_stats(p.tl, sb, N, mr, s_rs, s_st);
        }
      }
    }
  }
  __syncthreads();
  if (stats)
    for (int n = tid; n < 2 * N; n += blockDim.x)
      if (s_st[n] != 0.f) atomicAdd(&st.sums_out[n], (double)s_st[n]);
}

// ------------------------------------------------------------------------------------------------------------------
// GCN layer backward (blockIdx.y = stack)
// ------------------------------------------------------------------------------------------------------------------
__host__ __device__ inline int bwd_groups(int ldi, int ldo) {
  int g = kThreads / ((ldi >> 2) * (ldo >> 2));
  return g < 1 ? 1 : (g > 2 ? 2 : g);
}
static size_t bwd_smem(const gp_pk_layer_bwd_args& p) {
  size_t best = 0;
  for (int s = 0; s < p.ns; ++s) {
    Count c;
    const int ldi = r4(p.s[s].in.d), ldo = r4(p.s[s].dout), mr = p.tl.max_rows;
    c.take(mr * ldo); c.take(mr * ldi); c.take(mr * ldi);
    c.take(ldi * ldo); c.take(ldo * ldi); c.take(bwd_groups(ldi, ldo) * ldi * ldo); c.take(ldo);
    for (int i = 0; i < 4; ++i) c.take(2 * p.N);
    c.take(p.N); c.take(mr); c.take(mr);
    for (int i = 0; i < 2; ++i) { c.take(2 * mr); c.take(2 * mr * kEll); }
    best = best > c.n ? best : c.n;
  }
  return best * sizeof(float);
}

__global__ void __launch_bounds__(kThreads, 3)
layer_bwd_kernel(const gp_pk_layer_bwd_args p) {
  extern __shared__ __align__(16) float sm[];
  const gp_pk_stack_bwd& st = p.s[blockIdx.y];
  const int N = p.N, B = p.tl.B, din = st.in.d, dout = st.dout, ldi = r4(din), ldo = r4(dout), mr = p.tl.max_rows;
  const int groups = bwd_groups(ldi, ldo);
  Carve cv{sm};
  float* bV = cv.take(mr * ldo);      // dV
  float* bH = cv.take(mr * ldi);      // Hin
  float* bU = cv.take(mr * ldi);      // U = A Hin, then dU = dV W^T
  float* s_w = cv.take(ldi * ldo);
  float* s_wt = cv.take(ldo * ldi);
  float* s_dw = cv.take(groups * ldi * ldo);
  float* s_db = cv.take(ldo);
  float* s_bni = cv.take(2 * N);
  float* s_bno = cv.take(2 * N);
  float* s_m = cv.take(2 * N);
  float* s_mp = cv.take(2 * N);
  float* s_scr = cv.take(N);
  int* s_gs = reinterpret_cast<int*>(cv.take(mr));
  int* s_gid = reinterpret_cast<int*>(cv.take(mr));
  int2* s_info = reinterpret_cast<int2*>(cv.take(2 * mr));
  int2* s_ell = reinterpret_cast<int2*>(cv.take(2 * mr * kEll));
  int2* s_info_in = reinterpret_cast<int2*>(cv.take(2 * mr));
  int2* s_ell_in = reinterpret_cast<int2*>(cv.take(2 * mr * kEll));
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  for (int idx = tid; idx < ldi * ldo; idx += blockDim.x) {
    const int k = idx / ldo, n = idx - k * ldo;
    s_w[idx] = (k < din && n < dout) ? st.W[(long long)k * dout + n] : 0.f;
  }
  for (int idx = tid; idx < ldo * ldi; idx += blockDim.x) {
    const int n = idx / ldi, k = idx - n * ldi;
    s_wt[idx] = (k < din && n < dout) ? st.W[(long long)k * dout + n] : 0.f;
  }
  for (int idx = tid; idx < groups * ldi * ldo; idx += blockDim.x) s_dw[idx] = 0.f;
  for (int n = tid; n < ldo; n += blockDim.x) s_db[n] = 0.f;
  for (int n = tid; n < 2 * N; n += blockDim.x) s_mp[n] = 0.f;
  bn_stats(st.in, p.cnt_pad, N, B, s_bni, s_bni + N);
  bn_stats(st.out, p.cnt_pad, N, B, s_bno, s_bno + N);
  const bool bn = st.out.sums != nullptr;
  if (bn) {
    const double cnt = (double)B * (double)dout;
    for (int n = tid; n < 2 * N; n += blockDim.x) s_m[n] = (float)(st.msums[n] / cnt);
  }
  __syncthreads();

  // pad rows of a BatchNorm'd layer with a bias: no upstream gradient, but the batch means reach them; cnt_pad[n]
  // copies of one vector per node index feed the bias gradient (packed_blueprint.stack_backward)
  if (blockIdx.x == 0 && bn && st.b && st.db && p.cnt_pad) {
    float nn = 0.f;
    for (int c = 0; c < dout; ++c) nn = fmaf(st.b[c], st.b[c], nn);
    const float rp = fmaxf(sqrtf(nn), kEpsNorm);
    const float *mean = s_bno, *istd = s_bno + N, *m1 = s_m, *m2 = s_m + N;
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
      s_db[c] += acc;
    }
    __syncthreads();
  }

  float dbacc[4] = {0.f, 0.f, 0.f, 0.f};               // bias gradient: columns lane, lane+32, ... of this warp's rows
  const int C4 = ldi >> 2, C4p = pow2_ge(C4);
  const int nsub = sub_count(p.tl);
  for (int run = blockIdx.x; run < nsub; run += gridDim.x) {
    {
      const Sub sb = sub_get(p.tl, run);
      const int r0 = sb.r0, nt = sb.nt, g0 = sb.g0, g1 = sb.g1;
      (void)g0; (void)g1;
      if (nt > 0) {
        __syncthreads();
        load_meta(p.tl, sb, s_gs, s_gid);
        stage_lists(p.adj, p.tl, sb, s_info, s_ell);
        if (st.need_dx && p.adj_in.info) stage_lists(p.adj_in, p.tl, sb, s_info_in, s_ell_in);
        __syncthreads();
        // ---- dV = d normalize . d(ReLU + BatchNorm) . gl, straight from global memory (one warp per row)
        for (int i = wid; i < nt; i += nw) {
          const int ni = i - s_gs[i], gid = s_gid[i];
          const float r = st.rnorm[r0 + i];
          float mean = 0.f, istd = 1.f, m1 = 0.f, m2 = 0.f;
          if (bn) { mean = s_bno[ni]; istd = s_bno[N + ni]; m1 = s_m[ni]; m2 = s_m[N + ni]; }
          float gv[4], yv[4];
          float pr = 0.f;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int c = lane + 32 * t;
            gv[t] = 0.f;
            yv[t] = 0.f;
            if (c < dout) {
              float dy = grad_at(st.gl, r0, i, gid, ni, c);
              const float y = st.out.y[(long long)(r0 + i) * st.out.ld + c];
              if (bn) {
                const float h = (fmaxf(y, 0.f) - mean) * istd;
                dy = y > 0.f ? (dy - m1 - h * m2) * istd : 0.f;
              }
              gv[t] = dy;
              yv[t] = y;
              pr = fmaf(y, dy, pr);
            }
          }
          pr = warp_sum(pr);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int c = lane + 32 * t;
            if (c < ldo) {
              const float dv = c < dout ? (r > kEpsNorm ? (gv[t] - yv[t] * pr) / r : gv[t] / kEpsNorm) : 0.f;
              bV[i * ldo + c] = dv;
              dbacc[t] += dv;
            }
          }
        }
        load_rows(st.in, r0, nt, s_gs, s_bni, s_bni + N, bH, ldi, 0, ldi);
        __syncthreads();
        gather4(p.adj, p.tl, sb, s_gs, s_gid, s_info, s_ell, bH, ldi, bU);
        __syncthreads();
        gemm_tn_acc(bU, ldi, bV, ldo, ldi >> 2, ldo >> 2, 0, nt, s_dw, ldo, groups);
        if (st.need_dx) {
          __syncthreads();
          {
            float* du = bU;
            gemm_nn(bV, ldo, s_wt, ldi, nt, ldi >> 2, ldo >> 2, [=](int m, int n, float a) { du[m * ldi + n] = a; });
          }
          __syncthreads();
          if (st.dadj) {
            const int nf = p.tl.nfix;
            for (int idx = tid; idx < nt * nf; idx += blockDim.x) {
              const int i = idx / nf, j = idx - i * nf;               // dA[g][ni][j] = <dU[i], Hin[gs + j]>
              const float* a = bU + i * ldi;
              const float* h = bH + (s_gs[i] + j) * ldi;
              float acc = 0.f;
              for (int c = 0; c < ldi; c += 4) {
                const float4 x = *reinterpret_cast<const float4*>(a + c);
                const float4 y = *reinterpret_cast<const float4*>(h + c);
                acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
              }
              float* o = st.dadj + ((long long)s_gid[i] * nf + (i - s_gs[i])) * nf + j;
              *o = st.dadj_acc ? *o + acc : acc;
            }
          }
          // gl_prev = gz_prev + A^T dU, and its two batch sums per node index
          const int total = nt * C4p;
          for (int base = 0; base < total; base += blockDim.x) {
            const bool act = base + tid < total;
            const int item = act ? base + tid : total - 1;
            const int i = item / C4p, q = item - i * C4p;
            const int gs = s_gs[i], gid = s_gid[i], ni = i - gs;
            float a1 = 0.f, a2 = 0.f;
            if (q < C4) {
              const float4 d4 = gather_row4(p.adj_in, p.tl, sb, i, gs, gid, s_info_in,
                                            p.adj_in.info ? s_ell_in : s_ell, bU, ldi, q);
              const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) {
                const int c = 4 * q + cc;
                if (c < din) {
                  const float g = dv[cc] + grad_at(st.gz_prev, r0, i, gid, ni, c);
                  if (act) st.gl_prev[(long long)(r0 + i) * din + c] = g;
                  a1 += g;
                  a2 = fmaf(g, bH[i * ldi + c], a2);
                }
              }
            }
            if (st.msums_prev) {
              for (int o = C4p >> 1; o > 0; o >>= 1) {
                a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                a2 += __shfl_xor_sync(0xffffffffu, a2, o);
              }
              if (act && q == 0) {
                atomicAdd(&s_mp[ni], a1);
                atomicAdd(&s_mp[N + ni], a2);
              }
            }
          }
        }
      }
    }
  }
  if (st.db) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int c = lane + 32 * t;
      if (c < dout && dbacc[t] != 0.f) atomicAdd(&s_db[c], dbacc[t]);
    }
  }
  __syncthreads();
  for (int idx = tid; idx < din * dout; idx += blockDim.x) {
    const int k = idx / dout, n = idx - k * dout;
    float acc = 0.f;
    for (int g = 0; g < groups; ++g) acc += s_dw[g * ldi * ldo + k * ldo + n];
    if (acc != 0.f) atomicAdd(&st.dW[idx], acc);
  }
  if (st.db)
    for (int c = tid; c < dout; c += blockDim.x)
      if (s_db[c] != 0.f) atomicAdd(&st.db[c], s_db[c]);
  if (st.msums_prev)
    for (int n = tid; n < 2 * N; n += blockDim.x)
      if (s_mp[n] != 0.f) atomicAdd(&st.msums_prev[n], (double)s_mp[n]);
}

// ------------------------------------------------------------------------------------------------------------------
// Fast path of the layer kernels for din <= 32 and dout <= 32 (every level-0 layer at ENZYMES shapes, the pooled
// level's layers after its first): ONE WARP PER ROW, lane = column.  The weight column W[:, lane] (forward) or the
// weight row W[lane, :] plus the dW column (backward) live in registers; the only shared-memory traffic per row is the
// neighbour rows of the gather and a 128-byte broadcast buffer.  ~80 warp instructions per row and stack instead of
// ~380 for the 4x4-tiled kernels (ncu, profiles/r2_ncu_packed.md): no index arithmetic, no staging of V.
// ------------------------------------------------------------------------------------------------------------------
static size_t fwd_row_smem(const gp_pk_layer_fwd_args& p) {
  size_t best = 0;
  for (int s = 0; s < p.ns; ++s) {
    Count c;
    const int ldi = r4(p.s[s].in.d), mr = p.tl.max_rows;
    c.take(mr * ldi); c.take((kThreads / 32) * 64); c.take(2 * p.N); c.take(2 * p.N); c.take(mr); c.take(mr);
    c.take(2 * mr); c.take(2 * mr * kEll); c.take(2 * mr);
    best = best > c.n ? best : c.n;
  }
  return best * sizeof(float);
}

__global__ void __launch_bounds__(kThreads, 3)
layer_fwd_row_kernel(const gp_pk_layer_fwd_args p) {
  extern __shared__ __align__(16) float sm[];
  const gp_pk_stack_fwd& st = p.s[blockIdx.y];
  const int N = p.N, din = st.in.d, dout = st.dout, ldi = r4(din), mr = p.tl.max_rows;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  Carve cv{sm};
  float* bH = cv.take(mr * ldi);
  float* su = cv.take((kThreads / 32) * 64) + wid * 64;
  float* s_bn = cv.take(2 * N);
  float* s_st = cv.take(2 * N);
  int* s_gs = reinterpret_cast<int*>(cv.take(mr));
  int* s_gid = reinterpret_cast<int*>(cv.take(mr));
  int2* s_info = reinterpret_cast<int2*>(cv.take(2 * mr));
  int2* s_ent = reinterpret_cast<int2*>(cv.take(2 * mr * kEll));
  float* s_rs = cv.take(2 * mr);
  float wreg[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) wreg[k] = (k < din && lane < dout) ? __ldg(st.W + (long long)k * dout + lane) : 0.f;
  const float bias = (st.b && lane < dout) ? __ldg(st.b + lane) : 0.f;
  for (int n = tid; n < 2 * N; n += blockDim.x) s_st[n] = 0.f;
  bn_stats(st.in, p.cnt_pad, N, p.tl.B, s_bn, s_bn + N);
  const bool stats = st.sums_out != nullptr;
  const int nsub = sub_count(p.tl);
  for (int run = blockIdx.x; run < nsub; run += gridDim.x) {
    {
      const Sub sb = sub_get(p.tl, run);
      const int r0 = sb.r0, nt = sb.nt, g0 = sb.g0, g1 = sb.g1;
      (void)g0; (void)g1;
      if (nt > 0) {
        __syncthreads();
        load_meta(p.tl, sb, s_gs, s_gid);
        stage_lists(p.adj, p.tl, sb, s_info, s_ent);
        __syncthreads();
        load_rows_warp(st.in, sb, s_gs, s_bn, s_bn + N, bH, ldi);
        __syncthreads();
        for (int ib = wid; ib < nt; ib += 2 * nw) {        // two rows per iteration: their chains interleave
          int ii[2], nis[2];
          float vv[2];
          const bool two = ib + nw < nt;
          ii[0] = ib;
          ii[1] = two ? ib + nw : ib;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int i = ii[r], gs = s_gs[i];
            nis[r] = i - gs;
            su[r * 32 + lane] = gather_lane(p.adj, p.tl, sb, i, gs, s_gid[i], nis[r], s_info, s_ent, bH, ldi, lane);
          }
          __syncwarp();
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            float a0 = bias, a1 = 0.f;
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4)
              if (4 * k4 < ldi) {
                const float4 uu = *reinterpret_cast<const float4*>(su + r * 32 + 4 * k4);
                a0 = fmaf(uu.x, wreg[4 * k4], a0);
                a1 = fmaf(uu.y, wreg[4 * k4 + 1], a1);
                a0 = fmaf(uu.z, wreg[4 * k4 + 2], a0);
                a1 = fmaf(uu.w, wreg[4 * k4 + 3], a1);
              }
            vv[r] = a0 + a1;                               // lanes >= dout: 0
          }
          __syncwarp();
          float ss[2], s1[2], s2[2];
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const float q = fmaxf(vv[r], 0.f);
            ss[r] = vv[r] * vv[r];
            s1[r] = q;
            s2[r] = q * q;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              ss[r] += __shfl_xor_sync(0xffffffffu, ss[r], o);
              s1[r] += __shfl_xor_sync(0xffffffffu, s1[r], o);
              s2[r] += __shfl_xor_sync(0xffffffffu, s2[r], o);
            }
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            if (r == 1 && !two) break;
            const int i = ii[r];
            const float rn = fmaxf(sqrtf(ss[r]), kEpsNorm);
            if (lane < dout) st.y[(long long)(r0 + i) * dout + lane] = vv[r] / rn;
            if (lane == 0) {
              st.rnorm[r0 + i] = rn;
              const float inv = 1.f / rn;                 // sum relu(y) = sum relu(v) / rn (rn > 0)
              s_rs[i] = s1[r] * inv;
              s_rs[mr + i] = s2[r] * inv * inv;
            }
          }
        }
        if (stats) {
          __syncthreads();
          reduce_row_stats(p.tl, sb, N, mr, s_rs, s_st);
        }
      }
    }
  }
  __syncthreads();
  if (stats)
    for (int n = tid; n < 2 * N; n += blockDim.x)
      if (s_st[n] != 0.f) atomicAdd(&st.sums_out[n], (double)s_st[n]);
}

// ------------------------------------------------------------------------------------------------------------------
// pooling forward / backward
// ------------------------------------------------------------------------------------------------------------------
constexpr int kPoolThreads = 512;
struct PoolDims {
  int ldF, ldFa, ldK, Fa, F, K, gw;
};
__host__ __device__ inline PoolDims pool_dims(const gp_pk_pool_args& p) {
  PoolDims d;
  d.F = p.z.F; d.Fa = p.za.F; d.K = p.K;
  d.ldF = r4(d.F); d.ldFa = r4(d.Fa); d.ldK = r4(d.K);
  int g = kPoolThreads / ((d.ldK >> 2) * (d.ldFa >> 2));
  d.gw = g < 1 ? 1 : (g > 4 ? 4 : g);
  return d;
}
static size_t pool_smem(const gp_pk_pool_args& p, const PoolDims& d, bool bwd) {
  Count c;
  const int mr = p.tl.max_rows;
  c.take(mr * d.ldF); c.take(mr * d.ldFa); c.take(mr * d.ldK); c.take(mr * d.ldK); c.take(mr); c.take(mr);
  c.take((p.z.L + p.za.L) * 2 * p.N);
  for (int i = 0; i < (bwd ? 2 : 1); ++i) { c.take(2 * mr); c.take(2 * mr * kEll); }
  if (!bwd) {
    c.take(d.ldFa * d.ldK); c.take(d.ldK);
  } else {
    c.take(mr * d.ldK); c.take(mr * d.ldK); c.take(d.ldK * d.ldFa); c.take(d.gw * d.ldK * d.ldFa); c.take(d.ldK);
  }
  return c.n * sizeof(float);
}

__device__ void concat_stats(const gp_pk_concat& z, const float* cnt_pad, int N, int B, float* s_bn) {
  for (int l = 0; l < z.L; ++l) bn_stats(z.slot[l], cnt_pad, N, B, s_bn + l * 2 * N, s_bn + l * 2 * N + N);
}
__device__ void concat_load(const gp_pk_concat& z, int N, int r0, int nt, const int* s_gs, const float* s_bn,
                            float* dst, int ldd) {
  int off = 0;
  for (int l = 0; l < z.L; ++l) {
    const int d = z.slot[l].d;
    const int dz = (l == z.L - 1) ? ldd - off : d;
    load_rows(z.slot[l], r0, nt, s_gs, s_bn + l * 2 * N, s_bn + l * 2 * N + N, dst, ldd, off, dz);
    off += d;
  }
}

__global__ void __launch_bounds__(kPoolThreads)
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
  int2* s_info_in = reinterpret_cast<int2*>(cv.take(2 * mr));
  int2* s_ell_in = reinterpret_cast<int2*>(cv.take(2 * mr * kEll));
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
  const int nsub = sub_count(p.tl);
  for (int run = blockIdx.x; run < nsub; run += gridDim.x) {
    {
      const Sub sb = sub_get(p.tl, run);
      const int r0 = sb.r0, nt = sb.nt, g0 = sb.g0, g1 = sb.g1;
      (void)g0; (void)g1;
      __syncthreads();
      if (nt > 0) {                                   // graphs without rows still own (all-zero) outputs
        load_meta(p.tl, sb, s_gs, s_gid);
        stage_lists(p.adj_in, p.tl, sb, s_info_in, s_ell_in);
      }
      __syncthreads();
      if (nt > 0) {
        concat_load(p.z, N, r0, nt, s_gs, s_bn, bZ, ldF);
        concat_load(p.za, N, r0, nt, s_gs, s_bna, bZa, ldFa);
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
      if (nt > 0) gather4(p.adj_in, p.tl, sb, s_gs, s_gid, s_info_in, s_ell_in, bS, ldK, bT);    // (A^T S)[j][k] = T[k][j]
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
}

__global__ void __launch_bounds__(kPoolThreads)
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
  int2* s_info = reinterpret_cast<int2*>(cv.take(2 * mr));
  int2* s_ell = reinterpret_cast<int2*>(cv.take(2 * mr * kEll));
  int2* s_info_in = reinterpret_cast<int2*>(cv.take(2 * mr));
  int2* s_ell_in = reinterpret_cast<int2*>(cv.take(2 * mr * kEll));
  float* bAS = cv.take(mr * ldK);
  float* bD = cv.take(mr * ldK);
  float* s_wp = cv.take(ldK * ldFa);
  float* s_dwp = cv.take(dm.gw * ldK * ldFa);
  float* s_dbp = cv.take(ldK);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  for (int idx = tid; idx < ldK * ldFa; idx += blockDim.x) {
    const int k = idx / ldFa, f = idx - k * ldFa;
    s_wp[idx] = (k < K && f < Fa) ? p.Wp[(long long)k * Fa + f] : 0.f;
  }
  for (int idx = tid; idx < dm.gw * ldK * ldFa; idx += blockDim.x) s_dwp[idx] = 0.f;
  for (int k = tid; k < ldK; k += blockDim.x) s_dbp[k] = 0.f;
  concat_stats(p.z, p.cnt_pad, N, p.tl.B, s_bn);
  concat_stats(p.za, p.cnt_pad, N, p.tl.B, s_bna);
  float dbacc[4] = {0.f, 0.f, 0.f, 0.f};
  const int nsub = sub_count(p.tl);
  for (int run = blockIdx.x; run < nsub; run += gridDim.x) {
    {
      const Sub sb = sub_get(p.tl, run);
      const int r0 = sb.r0, nt = sb.nt, g0 = sb.g0, g1 = sb.g1;
      (void)g0; (void)g1;
      if (nt > 0) {
        __syncthreads();
        load_meta(p.tl, sb, s_gs, s_gid);
        stage_lists(p.adj, p.tl, sb, s_info, s_ell);
        stage_lists(p.adj_in, p.tl, sb, s_info_in, s_ell_in);
        __syncthreads();
        concat_load(p.z, N, r0, nt, s_gs, s_bn, bZ, ldF);
        concat_load(p.za, N, r0, nt, s_gs, s_bna, bZa, ldFa);
        for (int idx = tid; idx < nt * ldK; idx += blockDim.x) {
          const int i = idx / ldK, k = idx - i * ldK;
          bS[idx] = k < K ? p.S[((long long)s_gid[i] * N + (i - s_gs[i])) * K + k] : 0.f;
        }
        __syncthreads();
        gather4(p.adj, p.tl, sb, s_gs, s_gid, s_info, s_ell, bS, ldK, bAS);
        gather4(p.adj_in, p.tl, sb, s_gs, s_gid, s_info_in, s_ell_in, bS, ldK, bAtS);
        // gz[i][f] = sum_k S[i][k] dX'[g][k][f] + readout scatter   (dX' read through L1: each element serves n_g rows;
        // one warp per row, lane = column, two rows in flight)
        for (int i0 = wid; i0 < nt; i0 += 2 * nw) {
          const int i1 = i0 + nw < nt ? i0 + nw : i0;
          const int ga = s_gid[i0], gb2 = s_gid[i1];
          const float* sa = bS + i0 * ldK;
          const float* sb2 = bS + i1 * ldK;
          for (int f = lane; f < F; f += 32) {
            const float* da = p.dxp + (long long)ga * K * F + f;
            const float* db = p.dxp + (long long)gb2 * K * F + f;
            float acc0 = 0.f, acc1 = 0.f;
            for (int k = 0; k < K; ++k) {
              acc0 = fmaf(sa[k], __ldg(da), acc0);
              acc1 = fmaf(sb2[k], __ldg(db), acc1);
              da += F;
              db += F;
            }
            const long long o0 = (long long)ga * p.ldo + f;
            if (p.arg[o0] == i0 - s_gs[i0]) acc0 += p.dout[o0];
            p.gz[(long long)(r0 + i0) * F + f] = acc0;
            if (i1 != i0) {
              const long long o1 = (long long)gb2 * p.ldo + f;
              if (p.arg[o1] == i1 - s_gs[i1]) acc1 += p.dout[o1];
              p.gz[(long long)(r0 + i1) * F + f] = acc1;
            }
          }
        }
        __syncthreads();
        // dS[i][k] = <Z[i], dX'[k]> + <AS[i], dA'[k]> + sum_k2 AtS[i][k2] dA'[k2][k] + dS_ext
        for (int idx = tid; idx < nt * ldK; idx += blockDim.x) {
          const int i = idx / ldK, k = idx - i * ldK;
          float acc = 0.f;
          if (k < K) {
            const int g = s_gid[i], il = i - s_gs[i];
            const float* z = bZ + i * ldF;
            const float* dx = p.dxp + ((long long)g * K + k) * F;
            for (int f = 0; f < F; ++f) acc = fmaf(z[f], __ldg(dx + f), acc);
            const float* as = bAS + i * ldK;
            const float* ats = bAtS + i * ldK;
            const float* da = p.dap + (long long)g * K * K;
            for (int k2 = 0; k2 < K; ++k2) {
              acc = fmaf(as[k2], __ldg(da + k * K + k2), acc);
              acc = fmaf(ats[k2], __ldg(da + k2 * K + k), acc);
            }
            if (p.dS_ext) acc += p.dS_ext[((long long)g * N + il) * K + k];
          }
          bD[idx] = acc;
        }
        __syncthreads();
        for (int i = wid; i < nt; i += nw) {                      // softmax backward: dT = s (dS - <dS, s>)
          float* d = bD + i * ldK;
          const float* s = bS + i * ldK;
          float pr = 0.f;
          for (int k = lane; k < K; k += 32) pr = fmaf(d[k], s[k], pr);
          pr = warp_sum(pr);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int k = lane + 32 * t;
            if (k < ldK) {
              const float v = k < K ? s[k] * (d[k] - pr) : 0.f;
              d[k] = v;
              dbacc[t] += v;
            }
          }
        }
        __syncthreads();
        gemm_tn_acc(bD, ldK, bZa, ldFa, ldK >> 2, ldFa >> 2, 0, nt, s_dwp, ldFa, dm.gw);
        {
          float* gza = p.gza;
          gemm_nn(bD, ldK, s_wp, ldFa, nt, ldFa >> 2, ldK >> 2, [=](int m, int n, float a) {
            if (n < Fa) gza[(long long)(r0 + m) * Fa + n] = a;
          });
        }
      }
    }
  }
  if (p.dbp) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int k = lane + 32 * t;
      if (k < K && dbacc[t] != 0.f) atomicAdd(&s_dbp[k], dbacc[t]);
    }
  }
  __syncthreads();
  for (int idx = tid; idx < K * Fa; idx += blockDim.x) {
    const int k = idx / Fa, f = idx - k * Fa;
    float v = 0.f;
    for (int g = 0; g < dm.gw; ++g) v += s_dwp[g * ldK * ldFa + k * ldFa + f];
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

static int grid_for(int smem_bytes, int want, int ny = 1, int reg_limit = 8) {
  // persistent CTAs, ONE wave: as many as fit per SM (228 KB of shared memory per SM, 1 KB reserved per CTA; at most
  // `reg_limit` by registers), shared between the `ny` slices of the grid's y dimension, at most `want`
  int per_sm = (int)(228 * 1024 / (smem_bytes + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > reg_limit ? reg_limit : per_sm);
  int g = (kNumSMs * per_sm) / ny;
  g = g < want ? g : want;
  return g < 1 ? 1 : g;
}

// dynamic shared memory opt-in, once per device and kernel (not a stream operation: safe under graph capture too)
#define GP_PK_SMEM(kernel, bytes)                                                                              \
  do {                                                                                                         \
    if ((bytes) > 226 * 1024)                                                                                  \
      return gp::fail(GP_ERR_UNSUPPORTED, "packed schedule: %zu bytes of shared memory", (size_t)(bytes));     \
    GP_CONFIG_ONCE(GP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024))); \
  } while (0)

static int tiles_upper(const gp_pk_tiling& t) {
  if (t.subs == nullptr) return (t.B + t.gpt - 1) / t.gpt;
  return t.gpt;            // ragged level: the caller passes an upper bound of the run count in gpt
}

}  // namespace pk
}  // namespace gp

using namespace gp;
using namespace gp::pk;

extern "C" int gp_pk_prepare(const int32_t* nb, int B, int N, int window, int max_rows, int32_t* rowptr, float* cnt_pad,
                             int32_t* subs, int32_t* meta, gp_stream_t stream) {
  GP_REQUIRE(rowptr && cnt_pad && subs && meta, "pk_prepare: null pointer");
  GP_REQUIRE(B > 0 && N > 0 && N <= kMaxN && window > 0 && max_rows >= N && max_rows >= window,
             "pk_prepare: B=%d N=%d (N <= %d), max_rows >= max(N, window)", B, N, kMaxN);
  GP_REQUIRE((reinterpret_cast<uintptr_t>(subs) & 15) == 0, "pk_prepare: subs must be 16-byte aligned");
  prepare_kernel<<<1, 1024, 0, S(stream)>>>(nb, B, N, window, max_rows, rowptr, cnt_pad, subs, meta);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_pk_build_lists(const float* adj, const int32_t* nb, const int32_t* rowptr, int B, int N,
                                 int32_t* info_out, int32_t* ell_out, int32_t* ovf_out, int32_t* info_in,
                                 int32_t* ell_in, int32_t* ovf_in, int32_t* cursors, long long capacity,
                                 int32_t* rowmeta, const float* x, int D, float* xpack, long long ldxp, const float* ax,
                                 int Da, float* axpack, long long ldaxp, gp_stream_t stream) {
  GP_REQUIRE(adj && rowptr && info_out && ell_out && ovf_out && info_in && ell_in && ovf_in && cursors && rowmeta,
             "pk_build_lists: null pointer");
  GP_REQUIRE(B > 0 && N > 0 && N <= kMaxN && capacity > 0 && capacity < (1ll << 31), "pk_build_lists: bad sizes");
  GP_REQUIRE((xpack == nullptr) || (x && D > 0 && ldxp >= D), "pk_build_lists: x / xpack");
  GP_REQUIRE((axpack == nullptr) || (ax && Da > 0 && ldaxp >= Da), "pk_build_lists: ax / axpack");
  GP_REQUIRE(((reinterpret_cast<uintptr_t>(ell_out) | reinterpret_cast<uintptr_t>(ell_in)) & 15) == 0,
             "pk_build_lists: ell arrays must be 16-byte aligned");
  const size_t smem = (size_t)(N * (N + 1) + 2 * (N + 1)) * sizeof(float);
  GP_PK_SMEM(build_lists_kernel, smem);
  const int grid = grid_for((int)smem, B);
  build_lists_kernel<<<grid, 128, smem, S(stream)>>>(
      adj, nb, rowptr, B, N, reinterpret_cast<int2*>(info_out), reinterpret_cast<int2*>(ell_out),
      reinterpret_cast<int2*>(ovf_out), reinterpret_cast<int2*>(info_in), reinterpret_cast<int2*>(ell_in),
      reinterpret_cast<int2*>(ovf_in), cursors, capacity, reinterpret_cast<int2*>(rowmeta), x, D, xpack, ldxp, ax, Da,
      axpack, ldaxp);
  GP_LAUNCHED();
  return GP_OK;
}

static int check_tiling(const gp_pk_tiling& t, int N) {
  GP_REQUIRE(t.B > 0 && t.max_rows > 0 && t.gpt > 0, "pk: bad tiling");
  GP_REQUIRE((t.rowptr != nullptr) == (t.subs != nullptr) && (t.rowptr != nullptr) == (t.nsub != nullptr) &&
             (t.rowptr != nullptr) == (t.rowmeta != nullptr), "pk: the ragged level needs rowptr, subs, nsub and rowmeta");
  GP_REQUIRE(t.rowptr != nullptr || (t.nfix > 0 && t.nfix <= N && t.gpt * t.nfix <= t.max_rows), "pk: uniform tiling");
  GP_REQUIRE(t.rowptr == nullptr || t.max_rows >= N, "pk: max_rows must hold one whole graph");
  GP_REQUIRE(N > 0 && N <= kMaxN, "pk: N = %d (<= %d)", N, kMaxN);
  return GP_OK;
}
static int check_adj(const gp_pk_adj& a, const gp_pk_tiling& t) {
  GP_REQUIRE((a.info && a.ell && a.entries) || (a.dense && t.rowptr == nullptr),
             "pk: adjacency (lists, or dense with a uniform tiling)");
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
  bool narrow = true;                       // every stack at most 32 wide: one warp per row, weights in registers
  for (int s = 0; s < a->ns; ++s) narrow = narrow && a->s[s].in.d <= 32 && a->s[s].dout <= 32;
  if (narrow) {
    const size_t smem = fwd_row_smem(*a);
    GP_PK_SMEM(layer_fwd_row_kernel, smem);
    const int g = grid_for((int)smem, tiles_upper(a->tl), a->ns, 3);
    layer_fwd_row_kernel<<<dim3(g, a->ns), kThreads, smem, S(stream)>>>(*a);
    GP_LAUNCHED();
    return GP_OK;
  }
  const size_t smem = fwd_smem(*a);
  GP_PK_SMEM(layer_fwd_kernel, smem);
  layer_fwd_kernel<<<dim3(grid_for((int)smem, tiles_upper(a->tl), a->ns, 3), a->ns), kThreads, smem, S(stream)>>>(*a);
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
  const size_t smem = bwd_smem(*a);
  GP_PK_SMEM(layer_bwd_kernel, smem);
  layer_bwd_kernel<<<dim3(grid_for((int)smem, tiles_upper(a->tl), a->ns, 3), a->ns), kThreads, smem, S(stream)>>>(*a);
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
  const size_t smem = pool_smem(*a, dm, false);
  GP_PK_SMEM(pool_fwd_kernel, smem);
  pool_fwd_kernel<<<grid_for((int)smem, tiles_upper(a->tl), 1, 1), kPoolThreads, smem, S(stream)>>>(*a);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_pk_pool_bwd(const gp_pk_pool_args* a, gp_stream_t stream) {
  GP_TRY(check_pool(a, true));
  const PoolDims dm = pool_dims(*a);
  const size_t smem = pool_smem(*a, dm, true);
  GP_PK_SMEM(pool_bwd_kernel, smem);
  pool_bwd_kernel<<<grid_for((int)smem, tiles_upper(a->tl), 1, 1), kPoolThreads, smem, S(stream)>>>(*a);
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
