// fp32 -> bf16 operand conversion (strided rows, optional zero padding of the row tail up to ld_out): the bf16 shadow
// copies of the tensor-core schedule that no GEMM / row-kernel epilogue produces (inputs, weights).
#include <cuda_bf16.h>
#include "common.cuh"

namespace gp {

__global__ void cvt_bf16_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y,
                                long long ldy, long long rows, int cols, int cols_pad) {
  const int cpr = (cols_pad + 7) / 8;                       // 8-element chunks per row
  const long long total = rows * cpr;
  const bool vec = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (ldy % 8 == 0) &&
                   ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cpr;
    const int c0 = (int)(i - r * cpr) * 8;
    const float* src = x + r * ldx + c0;
    __nv_bfloat16* dst = y + r * ldy + c0;
    if (vec && c0 + 8 <= cols) {
      const float4 a = *reinterpret_cast<const float4*>(src);
      const float4 c = *reinterpret_cast<const float4*>(src + 4);
      __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(c.x, c.y), h3 = __floats2bfloat162_rn(c.z, c.w);
      uint4 o;
      o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
      o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(dst) = o;
    } else {
      for (int j = 0; j < 8 && c0 + j < cols_pad; ++j)
        dst[j] = __float2bfloat16_rn(c0 + j < cols ? src[j] : 0.f);
    }
  }
}

}  // namespace gp

using namespace gp;

extern "C" int gp_cvt_f32_bf16(const float* x, long long ldx, void* y, long long ldy, long long rows, int cols,
                               int cols_pad, gp_stream_t stream) {
  GP_REQUIRE(x && y && rows > 0 && cols > 0 && cols_pad >= cols && ldy >= cols_pad && ldx >= cols, "cvt_bf16: bad args");
  const long long total = rows * ((cols_pad + 7) / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  cvt_bf16_kernel<<<(int)blocks, 256, 0, S(stream)>>>(x, ldx, reinterpret_cast<__nv_bfloat16*>(y), ldy, rows, cols,
                                                      cols_pad);
  GP_LAUNCHED();
  return GP_OK;
}
