// Row kernels of the tensor-core schedule that also emit the bf16 operand copy of their result, so no separate
// conversion pass touches HBM (all HBM-bound, one warp per row, 16-byte accesses when the row allows):
//   bn_finalize      per-node BatchNorm statistics from the per-row sums the GraphConv GEMM epilogue produced
//   bn_apply         H = (relu(Y) - mean_n) * invstd_n  -> fp32 concat slot + bf16 operand   (encoders.py:1062-1064)
//   bias_normalize_x V + b -> Y = V / max(||V||, eps)    -> fp32 in place + bf16 operand      (encoders.py:323-326)
//   softmax_mask_*_x masked assignment softmax fwd / bwd -> fp32 + bf16 operand               (encoders.py:1273-1275)
#include <cuda_bf16.h>
#include "common.cuh"

namespace gp {

constexpr float kEpsBnX = 1e-5f;
constexpr float kEpsNormX = 1e-12f;

__device__ __forceinline__ uint32_t packx(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
static inline bool al16x(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline bool al8x(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }
static inline int row_grid(long long rows) {
  long long blocks = (rows + 7) / 8;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  return (int)(blocks < 1 ? 1 : blocks);
}

// one warp per node index: mean / invstd over (batch, feature) from rowstat[b*N + n] = (sum x, sum x^2)
__global__ void bn_finalize_kernel(const float2* __restrict__ rowstat, int B, int N, int d, float* __restrict__ mean,
                                   float* __restrict__ invstd) {
  const int lane = threadIdx.x & 31;
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= N) return;
  float s1 = 0.f, s2 = 0.f;
  for (int b = lane; b < B; b += 32) {
    const float2 t = rowstat[(long long)b * N + n];
    s1 += t.x; s2 += t.y;
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if (lane == 0) {
    const double cnt = (double)B * (double)d;
    const double mu = (double)s1 / cnt;
    double var = (double)s2 / cnt - mu * mu;
    if (var < 0.0) var = 0.0;
    mean[n] = (float)mu;
    invstd[n] = (float)(1.0 / sqrt(var + (double)kEpsBnX));
  }
}

template <bool VEC>
__global__ void bn_apply_kernel(const float* __restrict__ y, long long ldy, const float* __restrict__ mean,
                                const float* __restrict__ invstd, long long rows, int N, int d, int relu, int bn,
                                float* __restrict__ h, long long ldh, __nv_bfloat16* __restrict__ hb, long long ldhb,
                                __nv_bfloat16* __restrict__ hb2, long long ldhb2) {
  const int lane = threadIdx.x & 31;
  const long long gw = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = gw; r < rows; r += nw) {
    int b_, n;
    row_split(r, N, b_, n);
    const float mu = bn ? mean[n] : 0.f, is = bn ? invstd[n] : 1.f;
    if (VEC) {
      for (int c = lane * 4; c < d; c += 128) {
        float4 v = *reinterpret_cast<const float4*>(y + r * ldy + c);
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        v.x = (v.x - mu) * is; v.y = (v.y - mu) * is; v.z = (v.z - mu) * is; v.w = (v.w - mu) * is;
        if (h != nullptr) *reinterpret_cast<float4*>(h + r * ldh + c) = v;
        const uint2 pk = make_uint2(packx(v.x, v.y), packx(v.z, v.w));
        if (hb != nullptr) *reinterpret_cast<uint2*>(hb + r * ldhb + c) = pk;
        if (hb2 != nullptr) *reinterpret_cast<uint2*>(hb2 + r * ldhb2 + c) = pk;
      }
    } else {
      for (int c = lane; c < d; c += 32) {
        float v = y[r * ldy + c];
        if (relu) v = fmaxf(v, 0.f);
        v = (v - mu) * is;
        if (h != nullptr) h[r * ldh + c] = v;
        if (hb != nullptr) hb[r * ldhb + c] = __float2bfloat16_rn(v);
        if (hb2 != nullptr) hb2[r * ldhb2 + c] = __float2bfloat16_rn(v);
      }
    }
  }
}

// V (+bias) -> Y = V / max(||V||, eps) in place, rnorm, optional bf16 copy; the row stays in registers (d <= 1024)
template <int VPL, int MINB = 1>
__global__ void __launch_bounds__(256, MINB) bias_normalize_x_kernel(float* __restrict__ v, const float* __restrict__ bias, float* __restrict__ rnorm,
                                        long long rows, int d, long long ld, int normalize,
                                        __nv_bfloat16* __restrict__ yb, long long ldyb) {
  const int lane = threadIdx.x & 31;
  const long long gw = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = gw; r < rows; r += nw) {
    float4 x[VPL];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = (lane + 32 * k) * 4;
      x[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < d) {
        x[k] = *reinterpret_cast<const float4*>(v + r * ld + c);
        if (bias != nullptr) {
          const float4 b = *reinterpret_cast<const float4*>(bias + c);
          x[k].x += b.x; x[k].y += b.y; x[k].z += b.z; x[k].w += b.w;
        }
        ss = fmaf(x[k].x, x[k].x, fmaf(x[k].y, x[k].y, fmaf(x[k].z, x[k].z, fmaf(x[k].w, x[k].w, ss))));
      }
    }
    float nrm = 1.f;
    if (normalize) {
      ss = warp_sum(ss);
      nrm = fmaxf(sqrtf(ss), kEpsNormX);
    }
    if (lane == 0 && rnorm != nullptr) rnorm[r] = nrm;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = (lane + 32 * k) * 4;
      if (c < d) {
        float4 o = x[k];
        if (normalize) { o.x /= nrm; o.y /= nrm; o.z /= nrm; o.w /= nrm; }
        *reinterpret_cast<float4*>(v + r * ld + c) = o;
        if (yb != nullptr) *reinterpret_cast<uint2*>(yb + r * ldyb + c) = make_uint2(packx(o.x, o.y), packx(o.z, o.w));
      }
    }
  }
}

// masked row softmax, in place, optional bf16 copy; the row stays in registers (K <= 1024, K % 4 == 0)
template <int VPL>
__global__ void softmax_fwd_x_kernel(float* __restrict__ t, const int32_t* __restrict__ nb, long long rows, int N,
                                     int K, __nv_bfloat16* __restrict__ sb, long long ldsb) {
  const int lane = threadIdx.x & 31;
  const long long gw = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = gw; r < rows; r += nw) {
    int b, n;
    row_split(r, N, b, n);
    const bool pad = nb != nullptr && n >= nb[b];
    float4 x[VPL];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = (lane + 32 * k) * 4;
      x[k] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      if (c < K && !pad) {
        x[k] = *reinterpret_cast<const float4*>(t + r * K + c);
        mx = fmaxf(mx, fmaxf(fmaxf(x[k].x, x[k].y), fmaxf(x[k].z, x[k].w)));
      }
    }
    float inv = 0.f;
    if (!pad) {
      mx = warp_max(mx);
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        x[k].x = expf(x[k].x - mx); x[k].y = expf(x[k].y - mx); x[k].z = expf(x[k].z - mx); x[k].w = expf(x[k].w - mx);
        s += (x[k].x + x[k].y) + (x[k].z + x[k].w);
      }
      s = warp_sum(s);
      inv = 1.f / s;
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = (lane + 32 * k) * 4;
      if (c < K) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!pad) o = make_float4(x[k].x * inv, x[k].y * inv, x[k].z * inv, x[k].w * inv);
        *reinterpret_cast<float4*>(t + r * K + c) = o;
        if (sb != nullptr) *reinterpret_cast<uint2*>(sb + r * ldsb + c) = make_uint2(packx(o.x, o.y), packx(o.z, o.w));
      }
    }
  }
}

// dT = S * (dS - <dS, S>) on real rows, 0 on pad rows; fp32 and/or bf16 output; per-block partial column sums
template <int VPL>
__global__ void softmax_bwd_x_kernel(const float* __restrict__ s, const float* __restrict__ ds,
                                     const int32_t* __restrict__ nb, long long rows, int N, int K,
                                     float* __restrict__ dt, __nv_bfloat16* __restrict__ dtb, long long lddtb,
                                     float* __restrict__ part) {
  __shared__ __align__(16) float colacc[8][VPL * 128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * 8 + warp, nw = (long long)gridDim.x * 8;
  float4 cs[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) cs[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r = gw; r < rows; r += nw) {
    int b, n;
    row_split(r, N, b, n);
    const bool pad = nb != nullptr && n >= nb[b];
    float4 sv[VPL], gv[VPL];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = (lane + 32 * k) * 4;
      sv[k] = gv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < K && !pad) {
        sv[k] = *reinterpret_cast<const float4*>(s + r * K + c);
        gv[k] = *reinterpret_cast<const float4*>(ds + r * K + c);
        dot = fmaf(sv[k].x, gv[k].x, fmaf(sv[k].y, gv[k].y, fmaf(sv[k].z, gv[k].z, fmaf(sv[k].w, gv[k].w, dot))));
      }
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = (lane + 32 * k) * 4;
      if (c < K) {
        const float4 o = make_float4(sv[k].x * (gv[k].x - dot), sv[k].y * (gv[k].y - dot), sv[k].z * (gv[k].z - dot),
                                     sv[k].w * (gv[k].w - dot));
        if (dt != nullptr) *reinterpret_cast<float4*>(dt + r * K + c) = o;
        if (dtb != nullptr) *reinterpret_cast<uint2*>(dtb + r * lddtb + c) = make_uint2(packx(o.x, o.y), packx(o.z, o.w));
        cs[k].x += o.x; cs[k].y += o.y; cs[k].z += o.z; cs[k].w += o.w;
      }
    }
  }
  if (part != nullptr) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) *reinterpret_cast<float4*>(&colacc[warp][(lane + 32 * k) * 4]) = cs[k];
    __syncthreads();
    for (int c = threadIdx.x; c < K; c += blockDim.x) {
      float tsum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) tsum += colacc[w][c];
      part[(long long)blockIdx.x * K + c] = tsum;
    }
  }
}

// K up to 2048 (VPL = 8 / 16): the dot product in a first pass, S and dS read again (L1 / L2 hits) for the update;
// 4 warps per block, the lanes keep the column sums only
template <int VPL>
__global__ void __launch_bounds__(128) softmax_bwd_wide_kernel(const float* __restrict__ s, const float* __restrict__ ds,
                                                               const int32_t* __restrict__ nb, long long rows, int N,
                                                               int K, float* __restrict__ dt,
                                                               __nv_bfloat16* __restrict__ dtb, long long lddtb,
                                                               float* __restrict__ part) {
  __shared__ __align__(16) float colacc[4][VPL * 128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * 4 + warp, nw = (long long)gridDim.x * 4;
  float4 cs[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) cs[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r = gw; r < rows; r += nw) {
    int b, n;
    row_split(r, N, b, n);
    const bool pad = nb != nullptr && n >= nb[b];
    float dot = 0.f;
    if (!pad) {
#pragma unroll 4
      for (int k = 0; k < VPL; ++k) {
        const int c = (lane + 32 * k) * 4;
        if (c < K) {
          const float4 sv = *reinterpret_cast<const float4*>(s + r * K + c);
          const float4 gv = *reinterpret_cast<const float4*>(ds + r * K + c);
          dot = fmaf(sv.x, gv.x, fmaf(sv.y, gv.y, fmaf(sv.z, gv.z, fmaf(sv.w, gv.w, dot))));
        }
      }
      dot = warp_sum(dot);
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = (lane + 32 * k) * 4;
      if (c < K) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!pad) {
          const float4 sv = *reinterpret_cast<const float4*>(s + r * K + c);
          const float4 gv = *reinterpret_cast<const float4*>(ds + r * K + c);
          o = make_float4(sv.x * (gv.x - dot), sv.y * (gv.y - dot), sv.z * (gv.z - dot), sv.w * (gv.w - dot));
        }
        if (dt != nullptr) *reinterpret_cast<float4*>(dt + r * K + c) = o;
        if (dtb != nullptr) *reinterpret_cast<uint2*>(dtb + r * lddtb + c) = make_uint2(packx(o.x, o.y), packx(o.z, o.w));
        cs[k].x += o.x; cs[k].y += o.y; cs[k].z += o.z; cs[k].w += o.w;
      }
    }
  }
  if (part != nullptr) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) *reinterpret_cast<float4*>(&colacc[warp][(lane + 32 * k) * 4]) = cs[k];
    __syncthreads();
    for (int c = threadIdx.x; c < K; c += blockDim.x)
      part[(long long)blockIdx.x * K + c] = (colacc[0][c] + colacc[1][c]) + (colacc[2][c] + colacc[3][c]);
  }
}

int colsum(const float* x, long long rows, int d, long long ld, float* out, int accumulate, float* ws,
           cudaStream_t st);
int bias_normalize(float* v, const float* bias, float* rnorm, long long rows, int d, long long ld, int normalize,
                   cudaStream_t st);

}  // namespace gp

using namespace gp;

extern "C" int gp_bn_finalize(const float* rowstat, int B, int N, int d, float* mean, float* invstd,
                              gp_stream_t stream) {
  GP_REQUIRE(rowstat && mean && invstd && B > 0 && N > 0 && d > 0 && al8x(rowstat), "bn_finalize: bad args");
  const int blocks = (N * 32 + 255) / 256;
  bn_finalize_kernel<<<blocks, 256, 0, S(stream)>>>(reinterpret_cast<const float2*>(rowstat), B, N, d, mean, invstd);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_bn_apply(const float* y, long long ldy, const float* mean, const float* invstd, int B, int N, int d,
                           int relu, int bn, float* h, long long ldh, void* h_bf16, long long ldhb, void* h_bf16_2,
                           long long ldhb2, gp_stream_t stream) {
  GP_REQUIRE(y && (h || h_bf16) && B > 0 && N > 0 && d > 0 && ldy >= d, "bn_apply: bad args");
  GP_REQUIRE(!bn || (mean && invstd), "bn_apply: bn needs mean/invstd");
  const long long rows = (long long)B * N;
  GP_REQUIRE(rows <= 0x7fffffffLL, "B * N must stay below 2^31 rows");
  __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(h_bf16);
  __nv_bfloat16* hb2 = reinterpret_cast<__nv_bfloat16*>(h_bf16_2);
  const bool vec = d % 4 == 0 && al16x(y) && ldy % 4 == 0 && (!h || (al16x(h) && ldh % 4 == 0)) &&
                   (!hb || (al8x(hb) && ldhb % 4 == 0)) && (!hb2 || (al8x(hb2) && ldhb2 % 4 == 0));
  if (vec) bn_apply_kernel<true><<<row_grid(rows), 256, 0, S(stream)>>>(y, ldy, mean, invstd, rows, N, d, relu, bn, h, ldh, hb, ldhb, hb2, ldhb2);
  else     bn_apply_kernel<false><<<row_grid(rows), 256, 0, S(stream)>>>(y, ldy, mean, invstd, rows, N, d, relu, bn, h, ldh, hb, ldhb, hb2, ldhb2);
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_bias_normalize_x(float* v, const float* bias, float* rnorm, long long rows, int d, long long ld,
                                   int normalize, void* y_bf16, long long ldyb, gp_stream_t stream) {
  GP_REQUIRE(v && rows > 0 && d > 0 && ld >= d, "bias_normalize_x: bad args");
  GP_REQUIRE(!normalize || rnorm, "bias_normalize_x: normalize needs rnorm");
  __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(y_bf16);
  const bool vec = d % 4 == 0 && d <= 2048 && al16x(v) && ld % 4 == 0 && (!bias || al16x(bias)) &&
                   (!yb || (al8x(yb) && ldyb % 4 == 0));
  GP_REQUIRE(vec || yb == nullptr, "bias_normalize_x: the bf16 copy needs d %% 4 == 0, d <= 2048 and 16-byte aligned rows");
  if (!vec) return bias_normalize(v, bias, rnorm, rows, d, ld, normalize, S(stream));
  const int g = row_grid(rows);
  if (d <= 128)      bias_normalize_x_kernel<1><<<g, 256, 0, S(stream)>>>(v, bias, rnorm, rows, d, ld, normalize, yb, ldyb);
  else if (d <= 256) bias_normalize_x_kernel<2><<<g, 256, 0, S(stream)>>>(v, bias, rnorm, rows, d, ld, normalize, yb, ldyb);
  else if (d <= 512) bias_normalize_x_kernel<4><<<g, 256, 0, S(stream)>>>(v, bias, rnorm, rows, d, ld, normalize, yb, ldyb);
  else if (d <= 1024) bias_normalize_x_kernel<8><<<g, 256, 0, S(stream)>>>(v, bias, rnorm, rows, d, ld, normalize, yb, ldyb);
  else {
    static int wminb = -1;
    if (wminb < 0) { const char* e = getenv("GP_WIDE_MINB"); wminb = e != nullptr ? atoi(e) : 1; }   /* 80 registers: six / three blocks per SM, measured 2.31 -> 1.95 ms at cfg5 */
    if (wminb) bias_normalize_x_kernel<16, 3><<<g, 256, 0, S(stream)>>>(v, bias, rnorm, rows, d, ld, normalize, yb, ldyb);
    else       bias_normalize_x_kernel<16><<<g, 256, 0, S(stream)>>>(v, bias, rnorm, rows, d, ld, normalize, yb, ldyb);
  }
  GP_LAUNCHED();
  return GP_OK;
}

extern "C" int gp_softmax_mask_fwd_x(float* t, const int32_t* nb, int B, int N, int K, void* s_bf16, long long ldsb,
                                     gp_stream_t stream) {
  GP_REQUIRE(t && B > 0 && N > 0 && K > 0, "softmax_mask_fwd_x: bad args");
  __nv_bfloat16* sb = reinterpret_cast<__nv_bfloat16*>(s_bf16);
  const bool vec = K % 4 == 0 && K <= 2048 && al16x(t) && (!sb || (al8x(sb) && ldsb % 4 == 0));
  if (!vec) {
    GP_REQUIRE(sb == nullptr, "softmax_mask_fwd_x: the bf16 copy needs K %% 4 == 0 and K <= 2048");
    return gp_softmax_mask_fwd(t, nb, B, N, K, stream);
  }
  const long long rows = (long long)B * N;
  GP_REQUIRE(rows <= 0x7fffffffLL, "B * N must stay below 2^31 rows");
  const int g = row_grid(rows);
  if (K <= 128)      softmax_fwd_x_kernel<1><<<g, 256, 0, S(stream)>>>(t, nb, rows, N, K, sb, ldsb);
  else if (K <= 256) softmax_fwd_x_kernel<2><<<g, 256, 0, S(stream)>>>(t, nb, rows, N, K, sb, ldsb);
  else if (K <= 512) softmax_fwd_x_kernel<4><<<g, 256, 0, S(stream)>>>(t, nb, rows, N, K, sb, ldsb);
  else if (K <= 1024) softmax_fwd_x_kernel<8><<<g, 256, 0, S(stream)>>>(t, nb, rows, N, K, sb, ldsb);
  else               softmax_fwd_x_kernel<16><<<g, 256, 0, S(stream)>>>(t, nb, rows, N, K, sb, ldsb);
  GP_LAUNCHED();
  return GP_OK;
}

/* dcol (optional): column sums of dT (= the assign_pred bias gradient); ws >= (148*16 + 256) * K floats */
extern "C" int gp_softmax_mask_bwd_x(const float* s, const float* ds, const int32_t* nb, int B, int N, int K,
                                     float* dt, void* dt_bf16, long long lddtb, float* dcol, float* ws,
                                     gp_stream_t stream) {
  GP_REQUIRE(s && ds && (dt || dt_bf16) && B > 0 && N > 0 && K > 0, "softmax_mask_bwd_x: bad args");
  GP_REQUIRE(!dcol || ws, "softmax_mask_bwd_x: dcol needs ws");
  __nv_bfloat16* dtb = reinterpret_cast<__nv_bfloat16*>(dt_bf16);
  const bool vec = K % 4 == 0 && K <= 2048 && al16x(s) && al16x(ds) && (!dt || al16x(dt)) &&
                   (!dtb || (al8x(dtb) && lddtb % 4 == 0));
  const long long rows = (long long)B * N;
  GP_REQUIRE(rows <= 0x7fffffffLL, "B * N must stay below 2^31 rows");
  if (!vec) {
    GP_REQUIRE(dtb == nullptr && dt != nullptr, "softmax_mask_bwd_x: the bf16 copy needs K %% 4 == 0 and K <= 2048");
    GP_TRY(gp_softmax_mask_bwd(s, ds, nb, B, N, K, dt, stream));
    if (dcol) GP_TRY(colsum(dt, rows, K, K, dcol, 0, ws, S(stream)));
    return GP_OK;
  }
  const int g = row_grid(rows);
  float* part = dcol ? ws : nullptr;
  if (K <= 128)      softmax_bwd_x_kernel<1><<<g, 256, 0, S(stream)>>>(s, ds, nb, rows, N, K, dt, dtb, lddtb, part);
  else if (K <= 256) softmax_bwd_x_kernel<2><<<g, 256, 0, S(stream)>>>(s, ds, nb, rows, N, K, dt, dtb, lddtb, part);
  else if (K <= 512) softmax_bwd_x_kernel<4><<<g, 256, 0, S(stream)>>>(s, ds, nb, rows, N, K, dt, dtb, lddtb, part);
  else if (K <= 1024) softmax_bwd_wide_kernel<8><<<g, 128, 0, S(stream)>>>(s, ds, nb, rows, N, K, dt, dtb, lddtb, part);
  else               softmax_bwd_wide_kernel<16><<<g, 128, 0, S(stream)>>>(s, ds, nb, rows, N, K, dt, dtb, lddtb, part);
  GP_LAUNCHED();
  if (dcol) GP_TRY(colsum(ws, g, K, K, dcol, 0, ws + (long long)g * K, S(stream)));
  return GP_OK;
}
