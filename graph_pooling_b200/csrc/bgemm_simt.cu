// Generic strided batched fp32 GEMM on the FFMA pipe (gp_bgemm_f32).
//
// This is the parity anchor (GP_F32) for every dense contraction on the DiffPool path
// (torch.matmul / bmm / nn.Linear at encoders.py:319,322,1278,1279,1311,1024-1031) and the
// fallback for shapes the tcgen05 path does not take (D=3, K=10, ...).  Transposes are expressed
// through strides; per-graph limits (`lim`) clip M/N/K so that all-zero tiles beyond a graph's
// node count are skipped (padding-aware schedule).
//
// Tiling: BMxBNxBK block tile, 256 threads, TMxTN register micro-tile split into two half-tiles
// BM/2 (BN/2) apart so that the shared-memory reads of a warp are contiguous float4 / float2
// (conflict-free).  Global loads are register-staged (prefetch of tile k+1 overlaps the FFMAs of
// tile k) and the thread->element map is picked per operand so that the contiguous global
// dimension is the one the warp walks (coalesced for both row- and column-major operands).
#include "common.cuh"

namespace gp {

// COMP: compensated accumulation for long contractions.  Each BK-term tile is summed into a fresh partial, and the
// partials are added to the running sum with Kahan's correction, so the rounding error of a K-term product no longer
// grows like sqrt(K) * eps (1.2e-6 relative at K = 5000, measured) but stays at the ~1e-7 of a 16-term sum: what
// keeps the cancellation-prone bias gradients of the fp32 parity mode inside the 1e-5 rule at N = 2048 .. 5000.
template <int BM, int BN, int BK, int TM, int TN, bool COMP>
__global__ void __launch_bounds__(256)
bgemm_kernel(const gp_gemm g) {
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
  constexpr int EA = BM * BK / 256, EB = BN * BK / 256;
  constexpr int HM = TM / 2, HN = TN / 2;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int split = g.split_k > 1 ? g.split_k : 1;
  const int b = blockIdx.z / split, ks = blockIdx.z % split;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  int Me = g.M, Ne = g.N, Ke = g.K;
  if (g.lim != nullptr) {
    const int l = g.lim[b];
    if (g.lim_m) Me = min(Me, l);
    if (g.lim_n) Ne = min(Ne, l);
    if (g.lim_k) Ke = min(Ke, l);
  }
  float alpha = g.alpha;
  if (g.alpha_dev != nullptr) alpha *= *g.alpha_dev;

  float acc[TM][TN];
  float comp[COMP ? TM : 1][COMP ? TN : 1];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      acc[i][j] = 0.f;
      if (COMP) comp[COMP ? i : 0][COMP ? j : 0] = 0.f;
    }

  const bool live = (m0 < Me) && (n0 < Ne) && (Ke > 0);
  if (!live && g.beta == 1.f && g.bias == nullptr) return;   // nothing to add

  if (live) {
    const int ktiles = (Ke + BK - 1) / BK;
    const int per = (ktiles + split - 1) / split;
    const int kt0 = ks * per, kt1 = min(ktiles, kt0 + per);
    const float* Ab = g.A + (long long)b * g.sAb;
    const float* Bb = g.B + (long long)b * g.sBb;
    const bool a_mcontig = (g.sAm == 1 && g.sAk != 1);
    const bool b_kcontig = (g.sBk == 1 && g.sBn != 1);

    float ra[EA], rb[EB];
    auto load = [&](int kt) {
      const int k0 = kt * BK;
#pragma unroll
      for (int i = 0; i < EA; ++i) {
        const int e = tid + i * 256;
        int m, k;
        if (a_mcontig) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
        const int gm = m0 + m, gk = k0 + k;
        ra[i] = (gm < Me && gk < Ke) ? __ldg(Ab + (long long)gm * g.sAm + (long long)gk * g.sAk) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < EB; ++i) {
        const int e = tid + i * 256;
        int n, k;
        if (b_kcontig) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
        const int gn = n0 + n, gk = k0 + k;
        rb[i] = (gn < Ne && gk < Ke) ? __ldg(Bb + (long long)gk * g.sBk + (long long)gn * g.sBn) : 0.f;
      }
    };
    auto store = [&]() {
#pragma unroll
      for (int i = 0; i < EA; ++i) {
        const int e = tid + i * 256;
        int m, k;
        if (a_mcontig) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
        As[k][m] = ra[i];
      }
#pragma unroll
      for (int i = 0; i < EB; ++i) {
        const int e = tid + i * 256;
        int n, k;
        if (b_kcontig) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
        Bs[k][n] = rb[i];
      }
    };

    if (kt0 < kt1) {
      load(kt0);
      store();
      __syncthreads();
      for (int kt = kt0; kt < kt1; ++kt) {
        if (kt + 1 < kt1) load(kt + 1);
        float part[COMP ? TM : 1][COMP ? TN : 1];
        if (COMP) {
#pragma unroll
          for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) part[COMP ? i : 0][COMP ? j : 0] = 0.f;
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
          float a[TM], bb[TN];
#pragma unroll
          for (int i = 0; i < HM; ++i) {
            a[i] = As[kk][ty * HM + i];
            a[HM + i] = As[kk][BM / 2 + ty * HM + i];
          }
#pragma unroll
          for (int j = 0; j < HN; ++j) {
            bb[j] = Bs[kk][tx * HN + j];
            bb[HN + j] = Bs[kk][BN / 2 + tx * HN + j];
          }
#pragma unroll
          for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) {
              if (COMP) part[COMP ? i : 0][COMP ? j : 0] = fmaf(a[i], bb[j], part[COMP ? i : 0][COMP ? j : 0]);
              else acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
            }
        }
        if (COMP) {                                      // Kahan: acc += part, carrying the rounding error in comp
#pragma unroll
          for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) {
              const float y = part[COMP ? i : 0][COMP ? j : 0] - comp[COMP ? i : 0][COMP ? j : 0];
              const float t = acc[i][j] + y;
              comp[COMP ? i : 0][COMP ? j : 0] = (t - acc[i][j]) - y;
              acc[i][j] = t;
            }
        }
        __syncthreads();
        if (kt + 1 < kt1) { store(); __syncthreads(); }
      }
    }
  }

  // epilogue
  float* Cb = g.C + (long long)b * g.sCb;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + (i < HM ? ty * HM + i : BM / 2 + ty * HM + (i - HM));
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + (j < HN ? tx * HN + j : BN / 2 + tx * HN + (j - HN));
      if (n >= g.N) continue;
      float* c = Cb + (long long)m * g.sCm + (long long)n * g.sCn;
      const bool inside = (m < Me) && (n < Ne);
      float v = inside ? alpha * acc[i][j] : 0.f;
      if (split > 1) {
        if (inside && v != 0.f) atomicAdd(c, v);
      } else {
        if (g.bias != nullptr) v += g.bias[n];
        if (g.relu) v = fmaxf(v, 0.f);
        if (g.beta != 0.f) v += g.beta * (*c);
        *c = v;
      }
    }
  }
}

__global__ void scale_fill_kernel(float* c, long long sCb, long long sCm, long long sCn, int M, int N,
                                  int batch, float beta) {
  const long long total = (long long)batch * M * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const long long r = i / N;
    const int m = (int)(r % M);
    const long long b = r / M;
    float* p = c + b * sCb + (long long)m * sCm + (long long)n * sCn;
    *p = beta == 0.f ? 0.f : beta * (*p);
  }
}

template <int BM, int BN, int BK, int TM, int TN, bool COMP = false>
static int launch(const gp_gemm& g, cudaStream_t st) {
  const int split = g.split_k > 1 ? g.split_k : 1;
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, g.batch * split);
  bgemm_kernel<BM, BN, BK, TM, TN, COMP><<<grid, 256, 0, st>>>(g);
  GP_LAUNCHED();
  return GP_OK;
}

int bgemm_f32(const gp_gemm& g_in, cudaStream_t st) {
  gp_gemm g = g_in;
  GP_REQUIRE(g.A && g.B && g.C, "bgemm: null operand");
  GP_REQUIRE(g.M > 0 && g.N > 0 && g.K >= 0 && g.batch > 0, "bgemm: bad dims M=%d N=%d K=%d batch=%d",
             g.M, g.N, g.K, g.batch);
  GP_REQUIRE(g.batch * (long long)(g.split_k > 1 ? g.split_k : 1) <= 65535 * 1ll * 1024,
             "bgemm: batch too large");
  if (g.split_k > 1) {
    GP_REQUIRE(g.bias == nullptr && !g.relu, "bgemm: split_k with bias/relu");
    if (g.beta != 1.f) {
      const long long total = (long long)g.batch * g.M * g.N;
      int blocks = (int)((total + 255) / 256);
      if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
      scale_fill_kernel<<<blocks, 256, 0, st>>>(g.C, g.sCb, g.sCm, g.sCn, g.M, g.N, g.batch, g.beta);
      GP_LAUNCHED();
      g.beta = 1.f;
    }
  }
  // grid.z limit
  GP_REQUIRE((long long)g.batch * (g.split_k > 1 ? g.split_k : 1) <= 65535, "bgemm: batch*split_k > 65535");
  const long long tiles128 = (long long)((g.M + 127) / 128) * ((g.N + 127) / 128) * g.batch *
                             (g.split_k > 1 ? g.split_k : 1);
  // contraction length one accumulator sees: beyond 256 terms the compensated variants take over
  const int kchain = g.K / (g.split_k > 1 ? g.split_k : 1);
  if (kchain > 256) {
    if (g.M > 32 || g.N > 32) return launch<64, 64, 16, 4, 4, true>(g, st);
    return launch<32, 32, 16, 2, 2, true>(g, st);
  }
  if (g.M >= 96 && g.N >= 96 && tiles128 >= kNumSMs) return launch<128, 128, 16, 8, 8>(g, st);
  if (g.M > 32 || g.N > 32) return launch<64, 64, 16, 4, 4>(g, st);
  return launch<32, 32, 16, 2, 2>(g, st);
}

}  // namespace gp

extern "C" int gp_bgemm_f32(const gp_gemm* g, gp_stream_t stream) {
  if (g == nullptr) return gp::fail(GP_ERR_INVALID, "bgemm: null descriptor");
  return gp::bgemm_f32(*g, gp::S(stream));
}
