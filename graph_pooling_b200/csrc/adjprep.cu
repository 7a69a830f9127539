// Adjacency preparation for the tensor-core path (HBM-bound, one pass):
//   adj [B,N,N] (fp32 as train.py:197 feeds it, or uint8 {0,1} from a compact feed) -> bf16 operand [B,N,ld]
// and, in the same pass, two facts the later kernels exploit:
//   flags[0] != 0  <=>  some graph's adjacency is NOT symmetric
//   flags[1] != 0  <=>  some entry is outside {0,1}
// Each CTA owns a PAIR of mirrored 64x64 tiles (ti <= tj): both are read once (coalesced 16-byte loads),
// converted and written, and tile (ti,tj) is compared with the transpose of tile (tj,ti) through shared
// memory.  Tiles beyond a graph's node count are all-zero by the feed contract (graph_sampler.py:97-109):
// they are written as zeros without being read (padding-aware schedule).
#include <cuda_bf16.h>
#include <type_traits>
#include "common.cuh"

namespace gp {

constexpr int AT = 64;

template <typename T> struct Ld4;
template <> struct Ld4<float> {
  static __device__ __forceinline__ void ld(const float* p, bool vec, int nvalid, float (&v)[4]) {
    if (vec) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else { for (int e = 0; e < 4; ++e) v[e] = e < nvalid ? p[e] : 0.f; }
  }
};
template <> struct Ld4<uint8_t> {
  static __device__ __forceinline__ void ld(const uint8_t* p, bool vec, int nvalid, float (&v)[4]) {
    if (vec) { const uchar4 t = *reinterpret_cast<const uchar4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else { for (int e = 0; e < 4; ++e) v[e] = e < nvalid ? (float)p[e] : 0.f; }
  }
};

// bit-packed adjacency (gp_host_pack_adj_bits): bit (c & 7) of byte c >> 3 of the row; c is a multiple of 4 here
struct BitRow {};
template <> struct Ld4<BitRow> {
  static __device__ __forceinline__ void ld(const uint8_t* row, int c, int nvalid, float (&v)[4]) {
    const uint32_t byte = row[c >> 3] >> (c & 7);
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (e < nvalid && ((byte >> e) & 1u)) ? 1.f : 0.f;
  }
};

__device__ __forceinline__ uint32_t pk2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// loads tile rows r0.., cols c0.. of graph b into regs (4 passes of 16 rows, one float4 per thread), writes bf16
template <typename T> struct InPtr { using type = const T*; };
template <> struct InPtr<BitRow> { using type = const uint8_t*; };

template <typename T>
__device__ __forceinline__ void tile_io(typename InPtr<T>::type ab, long long ldin, __nv_bfloat16* __restrict__ ob,
                                        int N, long long ld, int r0, int c0, int nreal, bool vec_in, bool vec_out,
                                        float (&v)[4][4], int* non01) {
  const int tr = threadIdx.x >> 4, tc = (threadIdx.x & 15) * 4;
  const bool live = r0 < nreal && c0 < nreal;           // otherwise all-zero by contract: do not read
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + tr + 16 * i, c = c0 + tc;
#pragma unroll
    for (int e = 0; e < 4; ++e) v[i][e] = 0.f;
    if (live && r < N && c < N) {
      if constexpr (std::is_same<T, BitRow>::value) Ld4<T>::ld(ab + (long long)r * ldin, c, N - c, v[i]);
      else Ld4<T>::ld(ab + (long long)r * ldin + c, vec_in && c + 4 <= N, N - c, v[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + tr + 16 * i, c = c0 + tc;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (v[i][e] != 0.f && v[i][e] != 1.f) *non01 = 1;
    if (r < N && c < ld) {
      __nv_bfloat16* dst = ob + (long long)r * ld + c;
      if (vec_out && c + 4 <= ld) {
        *reinterpret_cast<uint2*>(dst) = make_uint2(pk2(v[i][0], v[i][1]), pk2(v[i][2], v[i][3]));
      } else {
        for (int e = 0; e < 4 && c + e < ld; ++e) dst[e] = __float2bfloat16_rn(v[i][e]);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
adj_prepare_kernel(typename InPtr<T>::type adj, long long ldin, const int32_t* __restrict__ nb, int N,
                   __nv_bfloat16* __restrict__ out, long long ld, int tiles, int* __restrict__ flags, bool vec_in,
                   bool vec_out) {
  __shared__ float X[AT][AT + 1];
  // blockIdx.x enumerates pairs (ti <= tj) row by row
  int ti = 0, rem = blockIdx.x;
  while (rem >= tiles - ti) { rem -= tiles - ti; ++ti; }
  const int tj = ti + rem;
  const int b = blockIdx.y;
  const int nreal = nb != nullptr ? min(nb[b], N) : N;
  typename InPtr<T>::type ab = adj + (long long)b * N * ldin;
  __nv_bfloat16* ob = out + (long long)b * N * ld;
  const int tr = threadIdx.x >> 4, tc = (threadIdx.x & 15) * 4;
  int non01 = 0, asym = 0;
  float v[4][4];
  tile_io<T>(ab, ldin, ob, N, ld, ti * AT, tj * AT, nreal, vec_in, vec_out, v, &non01);
  // the last tile column also owns the zero fill of the operand's padding columns [tiles*64, ld) -- none here:
  // ld <= round_up(N, 32) <= tiles * 64, so they fall inside tile tj == tiles-1 and tile_io's `c < ld` bound covers them.
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) X[tr + 16 * i][tc + e] = v[i][e];
  __syncthreads();
  if (tj != ti) {
    tile_io<T>(ab, ldin, ob, N, ld, tj * AT, ti * AT, nreal, vec_in, vec_out, v, &non01);
  }
  // compare (tj,ti)[r][c] with (ti,tj)[c][r]; on the diagonal the tile is compared with its own transpose
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (v[i][e] != X[tc + e][tr + 16 * i]) asym = 1;
  asym = __syncthreads_or(asym);
  non01 = __syncthreads_or(non01);
  if (threadIdx.x == 0 && flags != nullptr) {
    if (asym) atomicOr(&flags[0], 1);
    if (non01) atomicOr(&flags[1], 1);
  }
}

// out[b] = bf16( (*cond == 0) ? x[b] + x[b]^T : x[b] ), x [B,K,K] fp32 (the pooled-adjacency gradient dA').
// 32x32 tiles; the mirrored tile is read coalesced and transposed through shared memory.
__global__ void __launch_bounds__(256)
sym_select_kernel(const float* __restrict__ x, long long ldx, int K, const int32_t* __restrict__ cond,
                  __nv_bfloat16* __restrict__ out, long long ld) {
  __shared__ float Tt[32][33];
  const bool sym = cond != nullptr && *cond == 0;
  const float* xb = x + (long long)blockIdx.z * K * ldx;
  __nv_bfloat16* ob = out + (long long)blockIdx.z * K * ld;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (sym) {
    for (int r = ty; r < 32; r += 8) {                   // mirrored tile rows c0.., cols r0..
      const int gr = c0 + r, gc = r0 + tx;
      Tt[r][tx] = (gr < K && gc < K) ? xb[(long long)gr * ldx + gc] : 0.f;
    }
    __syncthreads();
  }
  for (int r = ty; r < 32; r += 8) {
    const int gr = r0 + r, gc = c0 + tx;
    if (gr >= K || gc >= ld) continue;
    float v = 0.f;
    if (gc < K) {
      v = xb[(long long)gr * ldx + gc];
      if (sym) v += Tt[tx][r];
    }
    ob[(long long)gr * ld + gc] = __float2bfloat16_rn(v);
  }
}

// Same, 64x64 tiles with 16-byte loads and 8-byte bf16 stores (needs K, ldx, ld multiples of 4 and aligned bases):
// the 32x32 scalar version moved its 0.4 GB at 1.8 TB/s.
__global__ void __launch_bounds__(256)
sym_select_vec_kernel(const float* __restrict__ x, long long ldx, int K, const int32_t* __restrict__ cond,
                      __nv_bfloat16* __restrict__ out, long long ld) {
  __shared__ float Tt[64][65];
  const bool sym = cond != nullptr && *cond == 0;
  const float* xb = x + (long long)blockIdx.z * K * ldx;
  __nv_bfloat16* ob = out + (long long)blockIdx.z * K * ld;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int q = threadIdx.x & 15, rr = threadIdx.x >> 4;     // 16 float4 per tile row, 16 rows per pass
  if (sym) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {                            // mirrored tile: rows c0.., cols r0..
      const int r = rr + 16 * i, gr = c0 + r, gc = r0 + q * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < K && gc < K) v = *reinterpret_cast<const float4*>(xb + (long long)gr * ldx + gc);
      Tt[r][q * 4] = v.x; Tt[r][q * 4 + 1] = v.y; Tt[r][q * 4 + 2] = v.z; Tt[r][q * 4 + 3] = v.w;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = rr + 16 * i, gr = r0 + r, gc = c0 + q * 4;
    if (gr >= K || gc >= ld) continue;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gc < K) {
      v = *reinterpret_cast<const float4*>(xb + (long long)gr * ldx + gc);
      if (sym) { v.x += Tt[q * 4][r]; v.y += Tt[q * 4 + 1][r]; v.z += Tt[q * 4 + 2][r]; v.w += Tt[q * 4 + 3][r]; }
    }
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(ob + (long long)gr * ld + gc) =
        make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
}

}  // namespace gp

using namespace gp;

extern "C" int gp_sym_select_bf16(const float* x, long long ldx, int B, int K, const int32_t* cond, void* out_bf16,
                                  long long ld, gp_stream_t stream) {
  GP_REQUIRE(x && out_bf16 && B > 0 && K > 0 && ld >= K && ldx >= K && B <= 65535, "sym_select_bf16: bad args");
  const bool vec = K % 4 == 0 && ldx % 4 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(out_bf16) & 7) == 0;
  if (vec)
    sym_select_vec_kernel<<<dim3((unsigned)((ld + 63) / 64), (unsigned)((K + 63) / 64), (unsigned)B), 256, 0, S(stream)>>>(
        x, ldx, K, cond, reinterpret_cast<__nv_bfloat16*>(out_bf16), ld);
  else
    sym_select_kernel<<<dim3((unsigned)((ld + 31) / 32), (unsigned)((K + 31) / 32), (unsigned)B), 256, 0, S(stream)>>>(
        x, ldx, K, cond, reinterpret_cast<__nv_bfloat16*>(out_bf16), ld);
  GP_LAUNCHED();
  return GP_OK;
}

namespace gp {
// Edge-list feed: the batch's adjacency never exists as fp32.  edges [E][2] are graph-local node ids,
// graph b owns edges eptr[b] .. eptr[b+1]; the operand slice is zero-filled by the caller (cudaMemsetAsync) and every
// edge writes 1.0 at (u, v) -- and at (v, u) for an undirected list.  Out-of-range ids are ignored.
template <class E2>          // int2 (32-bit ids) or ushort2 (16-bit ids)
__global__ void __launch_bounds__(256)
adj_from_edges_kernel(const E2* __restrict__ edges, const int32_t* __restrict__ eptr, int N,
                      __nv_bfloat16* __restrict__ out, long long ld, int undirected) {
  const int b = blockIdx.y;
  const int e0 = eptr[b], e1 = eptr[b + 1];
  __nv_bfloat16* ob = out + (long long)b * N * ld;
  const __nv_bfloat16 one = __float2bfloat16(1.0f);
  for (int e = e0 + blockIdx.x * blockDim.x + threadIdx.x; e < e1; e += gridDim.x * blockDim.x) {
    const E2 uv = edges[e];
    const int u = (int)uv.x, v = (int)uv.y;
    if ((unsigned)u >= (unsigned)N || (unsigned)v >= (unsigned)N) continue;
    ob[(long long)u * ld + v] = one;
    if (undirected) ob[(long long)v * ld + u] = one;
  }
}
__global__ void set_flags_kernel(int32_t* flags, int not_sym, int accumulate) {
  if (accumulate) { flags[0] |= not_sym; } else { flags[0] = not_sym; flags[1] = 0; }
}
}  // namespace gp

extern "C" int gp_adj_from_edges(const void* edges, int id_bytes, const int32_t* eptr, int B, int N,
                                 int max_edges_per_graph, int undirected, void* adj_bf16, long long ld, int32_t* flags,
                                 int accumulate_flags, gp_stream_t stream) {
  GP_REQUIRE(edges && eptr && adj_bf16 && B > 0 && N > 0 && ld >= N && ld - N < 32 && B <= 65535,
             "adj_from_edges: bad args (need N <= ld < N+32)");
  GP_REQUIRE(id_bytes == 4 || (id_bytes == 2 && N <= 65536), "adj_from_edges: node ids are 4 or 2 bytes (2: N <= 65536)");
  GP_REQUIRE((reinterpret_cast<uintptr_t>(edges) & (uintptr_t)(2 * id_bytes - 1)) == 0, "adj_from_edges: edge alignment");
  GP_CUDA(cudaMemsetAsync(adj_bf16, 0, (size_t)B * N * ld * sizeof(__nv_bfloat16), S(stream)));
  int gx = (max_edges_per_graph + 255) / 256;
  gx = gx < 1 ? 1 : (gx > 64 ? 64 : gx);
  if (id_bytes == 4)
    adj_from_edges_kernel<int2><<<dim3((unsigned)gx, (unsigned)B), 256, 0, S(stream)>>>(
        reinterpret_cast<const int2*>(edges), eptr, N, reinterpret_cast<__nv_bfloat16*>(adj_bf16), ld, undirected);
  else
    adj_from_edges_kernel<ushort2><<<dim3((unsigned)gx, (unsigned)B), 256, 0, S(stream)>>>(
        reinterpret_cast<const ushort2*>(edges), eptr, N, reinterpret_cast<__nv_bfloat16*>(adj_bf16), ld, undirected);
  GP_LAUNCHED();
  if (flags != nullptr) {
    set_flags_kernel<<<1, 1, 0, S(stream)>>>(flags, undirected ? 0 : 1, accumulate_flags);
    GP_LAUNCHED();
  }
  return GP_OK;
}

extern "C" int gp_adj_prepare(const void* adj, int adj_dtype, const int32_t* nb, int B, int N, void* adj_bf16,
                              long long ld, int32_t* flags, gp_stream_t stream) {
  return gp_adj_prepare_x(adj, adj_dtype, 0, nb, B, N, adj_bf16, ld, flags, 0, stream);
}

extern "C" int gp_adj_prepare_x(const void* adj, int adj_dtype, long long ld_in, const int32_t* nb, int B, int N,
                                void* adj_bf16, long long ld, int32_t* flags, int accumulate_flags,
                                gp_stream_t stream) {
  GP_REQUIRE(adj && adj_bf16 && B > 0 && N > 0 && ld >= N && ld - N < 32, "adj_prepare: bad args (need N <= ld < N+32)");
  GP_REQUIRE(adj_dtype >= 0 && adj_dtype <= 2, "adj_prepare: adj_dtype must be 0 (fp32), 1 (uint8) or 2 (bit-packed)");
  GP_REQUIRE(B <= 65535, "adj_prepare: B too large");
  if (ld_in <= 0) ld_in = adj_dtype == 2 ? (N + 7) / 8 : N;
  GP_REQUIRE(ld_in >= (adj_dtype == 2 ? (N + 7) / 8 : N), "adj_prepare: input row stride too small");
  const int tiles = (N + AT - 1) / AT;
  const long long pairs = (long long)tiles * (tiles + 1) / 2;
  GP_REQUIRE(pairs < (1LL << 31), "adj_prepare: N too large");
  if (flags != nullptr && !accumulate_flags) GP_CUDA(cudaMemsetAsync(flags, 0, 2 * sizeof(int32_t), S(stream)));
  dim3 grid((unsigned)pairs, (unsigned)B);
  const bool vec_out = (reinterpret_cast<uintptr_t>(adj_bf16) & 7) == 0 && ld % 4 == 0;
  __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(adj_bf16);
  if (adj_dtype == 0) {
    const bool vec_in = (reinterpret_cast<uintptr_t>(adj) & 15) == 0 && N % 4 == 0 && ld_in % 4 == 0;
    adj_prepare_kernel<float><<<grid, 256, 0, S(stream)>>>(reinterpret_cast<const float*>(adj), ld_in, nb, N, ob, ld,
                                                          tiles, flags, vec_in, vec_out);
  } else if (adj_dtype == 1) {
    const bool vec_in = (reinterpret_cast<uintptr_t>(adj) & 3) == 0 && N % 4 == 0 && ld_in % 4 == 0;
    adj_prepare_kernel<uint8_t><<<grid, 256, 0, S(stream)>>>(reinterpret_cast<const uint8_t*>(adj), ld_in, nb, N, ob, ld,
                                                            tiles, flags, vec_in, vec_out);
  } else {
    adj_prepare_kernel<BitRow><<<grid, 256, 0, S(stream)>>>(reinterpret_cast<const uint8_t*>(adj), ld_in, nb, N, ob, ld,
                                                           tiles, flags, false, vec_out);
  }
  GP_LAUNCHED();
  return GP_OK;
}
