// Set2Set readout of GcnSet2SetEncoder (method=base-set2set): /root/reference/set2set.py:33-57, called from
// encoders.py:1144-1157.  SURVEY.md 8(f) N4 -- outside the north-star path, built for API completeness.
//
// Per graph, with E [n, d] the (masked) node embeddings, a one-layer LSTM (input 2d, hidden d, gate order i f g o as
// torch.nn.LSTM) and q*_0 = 0, h_0 = c_0 = 0, the reference runs n sequential steps
//     (q_t, (h_t, c_t)) = LSTM(q*_{t-1}, (h_{t-1}, c_{t-1}))        q_t = h_t
//     e = E q_t ;  a = softmax over ALL n rows (pad rows have E = 0 -> e = 0, they keep weight) ;  r_t = a^T E
//     q*_t = [q_t, r_t]
// and returns q*_n.  The recurrence is strictly sequential per graph and independent across graphs: ONE CTA per
// graph walks the n steps with its state in shared memory; LSTM weights and E are read through L1 / L2 each step
// (a graph's E is 36 KB at ENZYMES sizes).  Everything the backward needs is written once per step:
//     QS [B, n+1, 2d]  q*_t for t = 0..n (row 0 = zeros, row n = the output)
//     G  [B, n, 4d]    gate activations (i, f, g, o)        CS [B, n, d]  cell state c_t
//     AT [B, n, n]     attention weights a_t
// The backward kernel walks the steps in reverse (BPTT) and emits per-step gradients only:
//     DZ [B, n+1, 4d]  gradient of the gate pre-activations (row n = zeros, so DZ rows pair with QS rows)
//     DR [B, n, d]     gradient of r_t                      DE [B, n, n]  gradient of the attention logits
// from which the host side forms, with the library's batched GEMM and column sums,
//     dW_ih = DZ^T QS ;  dW_hh = DZ^T QS[:, :d] ;  db_ih = db_hh = colsum(DZ)
//     dE_b  = AT_b^T DR_b + DE_b^T Q_b   (Q_b[t] = QS[b, t+1, :d]) , rows >= n_b zeroed (the mask of encoders.py:1080).
#include "common.cuh"

namespace gp {

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

// y[j] = dot(W[j, 0:cols], x) for the rows j owned by this warp (warp per row, lanes across the columns)
__device__ __forceinline__ float warp_dot(const float* __restrict__ w, const float* x, int cols, int lane) {
  float acc = 0.f;
  for (int k = lane; k < cols; k += 32) acc = fmaf(__ldg(w + k), x[k], acc);
  return warp_sum(acc);
}

struct S2sArgs {
  const float* E; long long ldE; const int32_t* nb; int N, d;
  const float* Wih; const float* Whh; const float* bih; const float* bhh;
  float* QS; float* G; float* CS; float* AT;             // forward: outputs; backward: inputs
  const float* dout; long long lddout;                    // backward: gradient of QS[:, n, :]
  float* DZ; float* DR; float* DE;                        // backward outputs
};

// shared-memory layout (floats): qs[2d] | c[d] | gate[4d] | ea[N] | part[8*d] | red[64]
__global__ void __launch_bounds__(256) set2set_fwd_kernel(const S2sArgs a) {
  extern __shared__ float sm[];
  const int d = a.d, N = a.N;
  float* qs = sm;
  float* c = qs + 2 * d;
  float* gate = c + d;
  float* ea = gate + 4 * d;
  float* part = ea + N;
  float* red = part + 8 * d;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nreal = a.nb != nullptr ? min(a.nb[b], N) : N;
  const float* Eb = a.E + (long long)b * N * a.ldE;
  float* QSb = a.QS + (long long)b * (N + 1) * 2 * d;
  float* Gb = a.G + (long long)b * N * 4 * d;
  float* CSb = a.CS + (long long)b * N * d;
  float* ATb = a.AT + (long long)b * N * N;
  for (int k = tid; k < 2 * d; k += 256) { qs[k] = 0.f; QSb[k] = 0.f; }
  for (int k = tid; k < d; k += 256) c[k] = 0.f;
  __syncthreads();
  for (int t = 0; t < N; ++t) {
    // 1. gate pre-activations: W_ih q* + b_ih + W_hh h + b_hh   (h = q*[0:d])
    for (int j = warp; j < 4 * d; j += 8) {
      float v = warp_dot(a.Wih + (long long)j * 2 * d, qs, 2 * d, lane) + warp_dot(a.Whh + (long long)j * d, qs, d, lane);
      if (lane == 0) gate[j] = v + (a.bih != nullptr ? a.bih[j] : 0.f) + (a.bhh != nullptr ? a.bhh[j] : 0.f);
    }
    __syncthreads();
    // 2. cell update; q = h
    for (int k = tid; k < d; k += 256) {
      const float gi = sigm(gate[k]), gf = sigm(gate[d + k]), gg = tanhf(gate[2 * d + k]), go = sigm(gate[3 * d + k]);
      const float cn = fmaf(gf, c[k], gi * gg);
      c[k] = cn;
      qs[k] = go * tanhf(cn);
      float* g4 = Gb + (long long)t * 4 * d;
      g4[k] = gi; g4[d + k] = gf; g4[2 * d + k] = gg; g4[3 * d + k] = go;
      CSb[(long long)t * d + k] = cn;
    }
    __syncthreads();
    // 3. attention logits over all N rows (pad rows: E = 0 -> e = 0)
    for (int j = warp; j < N; j += 8) {
      float v = 0.f;
      if (j < nreal) v = warp_dot(Eb + (long long)j * a.ldE, qs, d, lane);
      if (lane == 0) ea[j] = v;
    }
    __syncthreads();
    // 4. softmax over the N rows
    float mx = -INFINITY;
    for (int j = tid; j < N; j += 256) mx = fmaxf(mx, ea[j]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    float s = 0.f;
    for (int j = tid; j < N; j += 256) {
      const float p = expf(ea[j] - mx);
      ea[j] = p;
      s += p;
    }
    s = block_sum(s, red + 8);
    const float inv = 1.f / s;
    for (int j = tid; j < N; j += 256) {
      const float p = ea[j] * inv;
      ea[j] = p;
      ATb[(long long)t * N + j] = p;
    }
    __syncthreads();
    // 5. r = a^T E (real rows only): each warp takes a strided subset of rows, lanes across the features
    for (int k0 = 0; k0 < d; k0 += 32) {
      const int k = k0 + lane;
      float acc = 0.f;
      if (k < d)
        for (int j = warp; j < nreal; j += 8) acc = fmaf(ea[j], __ldg(Eb + (long long)j * a.ldE + k), acc);
      if (k < d) part[warp * d + k] = acc;
    }
    __syncthreads();
    for (int k = tid; k < d; k += 256) {
      float r = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) r += part[w * d + k];
      qs[d + k] = r;
    }
    __syncthreads();
    float* qn = QSb + (long long)(t + 1) * 2 * d;
    for (int k = tid; k < 2 * d; k += 256) qn[k] = qs[k];
  }
}

// shared-memory layout (floats): dqs[2d] | dc[d] | dz[4d] | ea[N] | part[8*d] | dr[d] | red[64]
__global__ void __launch_bounds__(256) set2set_bwd_kernel(const S2sArgs a) {
  extern __shared__ float sm[];
  const int d = a.d, N = a.N;
  float* dqs = sm;
  float* dc = dqs + 2 * d;
  float* dz = dc + d;
  float* ea = dz + 4 * d;
  float* part = ea + N;
  float* dr = part + 8 * d;
  float* red = dr + d;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nreal = a.nb != nullptr ? min(a.nb[b], N) : N;
  const float* Eb = a.E + (long long)b * N * a.ldE;
  const float* Gb = a.G + (long long)b * N * 4 * d;
  const float* CSb = a.CS + (long long)b * N * d;
  const float* ATb = a.AT + (long long)b * N * N;
  float* DZb = a.DZ + (long long)b * (N + 1) * 4 * d;
  float* DRb = a.DR + (long long)b * N * d;
  float* DEb = a.DE + (long long)b * N * N;
  for (int k = tid; k < 2 * d; k += 256) dqs[k] = a.dout[(long long)b * a.lddout + k];
  for (int k = tid; k < d; k += 256) dc[k] = 0.f;
  for (int k = tid; k < 4 * d; k += 256) DZb[(long long)N * 4 * d + k] = 0.f;
  __syncthreads();
  for (int t = N - 1; t >= 0; --t) {
    // (a) dr = dq*[d:2d]
    for (int k = tid; k < d; k += 256) { const float v = dqs[d + k]; dr[k] = v; DRb[(long long)t * d + k] = v; }
    __syncthreads();
    // (b) da_j = E_j . dr   (0 on pad rows) ;  s = sum_j a_j da_j
    float sp = 0.f;
    for (int j = warp; j < N; j += 8) {
      float v = 0.f;
      if (j < nreal) v = warp_dot(Eb + (long long)j * a.ldE, dr, d, lane);
      if (lane == 0) { ea[j] = v; sp = fmaf(ATb[(long long)t * N + j], v, sp); }
    }
    const float s = block_sum(sp, red);
    // (c) de_j = a_j (da_j - s)
    for (int j = tid; j < N; j += 256) {
      const float v = ATb[(long long)t * N + j] * (ea[j] - s);
      ea[j] = v;
      DEb[(long long)t * N + j] = v;
    }
    __syncthreads();
    // (d) dq += sum_j de_j E_j
    for (int k0 = 0; k0 < d; k0 += 32) {
      const int k = k0 + lane;
      float acc = 0.f;
      if (k < d)
        for (int j = warp; j < nreal; j += 8) acc = fmaf(ea[j], __ldg(Eb + (long long)j * a.ldE + k), acc);
      if (k < d) part[warp * d + k] = acc;
    }
    __syncthreads();
    // (e, f) LSTM cell backward -> dz (gate pre-activations), dc carried to step t-1
    for (int k = tid; k < d; k += 256) {
      float dh = dqs[k];
#pragma unroll
      for (int w = 0; w < 8; ++w) dh += part[w * d + k];
      const float* g4 = Gb + (long long)t * 4 * d;
      const float gi = g4[k], gf = g4[d + k], gg = g4[2 * d + k], go = g4[3 * d + k];
      const float ct = CSb[(long long)t * d + k];
      const float cp = t > 0 ? CSb[(long long)(t - 1) * d + k] : 0.f;
      const float tc = tanhf(ct);
      const float d_o = dh * tc;
      const float dct = fmaf(dh * go, 1.f - tc * tc, dc[k]);
      dc[k] = dct * gf;
      const float zi = dct * gg * gi * (1.f - gi);
      const float zf = dct * cp * gf * (1.f - gf);
      const float zg = dct * gi * (1.f - gg * gg);
      const float zo = d_o * go * (1.f - go);
      dz[k] = zi; dz[d + k] = zf; dz[2 * d + k] = zg; dz[3 * d + k] = zo;
      float* o4 = DZb + (long long)t * 4 * d;
      o4[k] = zi; o4[d + k] = zf; o4[2 * d + k] = zg; o4[3 * d + k] = zo;
    }
    __syncthreads();
    // (g) gradient of q*_{t-1}: W_ih^T dz (+ W_hh^T dz on the first d entries, h_{t-1} = q_{t-1}); thread per column
    for (int m = tid; m < 2 * d; m += 256) {
      float acc = 0.f;
      for (int j = 0; j < 4 * d; ++j) acc = fmaf(__ldg(a.Wih + (long long)j * 2 * d + m), dz[j], acc);
      if (m < d)
        for (int j = 0; j < 4 * d; ++j) acc = fmaf(__ldg(a.Whh + (long long)j * d + m), dz[j], acc);
      dqs[m] = acc;
    }
    __syncthreads();
  }
}

static size_t s2s_smem(int N, int d) { return (size_t)(16 * d + N + 64) * sizeof(float); }

}  // namespace gp

using namespace gp;

static int s2s_check(const float* E, long long ldE, int B, int N, int d, const float* Wih, const float* Whh,
                     size_t* smem) {
  GP_REQUIRE(E && Wih && Whh && B > 0 && N > 0 && d > 0 && ldE >= d, "set2set: bad args");
  GP_REQUIRE(B <= 2147483647 / 2, "set2set: batch too large");
  *smem = s2s_smem(N, d);
  GP_REQUIRE(*smem <= 200 * 1024, "set2set: n + 16 d floats of shared memory exceed 200 KB (n = %d, d = %d)", N, d);
  return GP_OK;
}

extern "C" int gp_set2set_fwd(const float* E, long long ldE, const int32_t* nb, int B, int N, int d,
                              const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                              float* qs, float* gates, float* cells, float* att, gp_stream_t stream) {
  size_t smem = 0;
  GP_TRY(s2s_check(E, ldE, B, N, d, w_ih, w_hh, &smem));
  GP_REQUIRE(qs && gates && cells && att, "set2set_fwd: null output");
  S2sArgs a = {};
  a.E = E; a.ldE = ldE; a.nb = nb; a.N = N; a.d = d;
  a.Wih = w_ih; a.Whh = w_hh; a.bih = b_ih; a.bhh = b_hh;
  a.QS = qs; a.G = gates; a.CS = cells; a.AT = att;
  if (smem > 48 * 1024) GP_CUDA(cudaFuncSetAttribute(set2set_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  set2set_fwd_kernel<<<B, 256, smem, S(stream)>>>(a);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_set2set_bwd(const float* E, long long ldE, const int32_t* nb, int B, int N, int d,
                              const float* w_ih, const float* w_hh, const float* gates, const float* cells,
                              const float* att, const float* dout, long long lddout, float* dz, float* dr, float* de,
                              gp_stream_t stream) {
  size_t smem = 0;
  GP_TRY(s2s_check(E, ldE, B, N, d, w_ih, w_hh, &smem));
  GP_REQUIRE(gates && cells && att && dout && dz && dr && de && lddout >= 2 * d, "set2set_bwd: bad args");
  S2sArgs a = {};
  a.E = E; a.ldE = ldE; a.nb = nb; a.N = N; a.d = d;
  a.Wih = w_ih; a.Whh = w_hh;
  a.G = const_cast<float*>(gates); a.CS = const_cast<float*>(cells); a.AT = const_cast<float*>(att);
  a.dout = dout; a.lddout = lddout; a.DZ = dz; a.DR = dr; a.DE = de;
  if (smem > 48 * 1024) GP_CUDA(cudaFuncSetAttribute(set2set_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  set2set_bwd_kernel<<<B, 256, smem, S(stream)>>>(a);
  GP_LAUNCHED();
  return GP_OK;
}
