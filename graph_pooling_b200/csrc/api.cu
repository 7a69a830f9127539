// C-ABI composites of the DiffPool path: GraphConv forward/backward and the pooling contractions,
// each a short fixed schedule of this library's own kernels on the caller's stream.
#include "common.cuh"

namespace gp {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int bgemm_f32(const gp_gemm& g, cudaStream_t st);
int bias_normalize(float* v, const float* bias, float* rnorm, long long rows, int d, long long ld,
                   int normalize, cudaStream_t st);
int colsum(const float* x, long long rows, int d, long long ld, float* out, int accumulate, float* ws,
           cudaStream_t st);

bool small_gcn_eligible(int B, int N, int din, int dout, int add_self);
bool small_pool_eligible(int B, int N, int K, int F);
int small_pool_fwd(const float* s, const float* z, long long ldz, const float* adj, const int32_t* nb, int B, int N,
                   int K, int F, float* xp, float* t, float* ap, cudaStream_t st);
int small_pool_bwd(const float* dxp, const float* dap, const float* s, const float* z, long long ldz, const float* adj,
                   const float* t, const int32_t* nb, int B, int N, int K, int F, float* dz, long long lddz, int acc_dz,
                   float* ds, int acc_ds, cudaStream_t st);
int small_gcn_fwd(const float* x, long long ldx, const float* adj, const float* w, const float* bias, const int32_t* nb,
                  int B, int N, int din, int dout, int normalize, float* u, float* y, long long ldy, float* rnorm,
                  cudaStream_t st);
int small_gcn_bwd(const float* dv, const float* u, const float* x, long long ldx, const float* adj, const float* w,
                  const int32_t* nb, int B, int N, int din, int dout, float* dw, float* db, float* dx, float* dadj,
                  float* ws, cudaStream_t st);

static gp_gemm mk(const float* A, const float* B, float* C, int M, int N, int K, int batch) {
  gp_gemm g;
  g.A = A; g.B = B; g.C = C; g.M = M; g.N = N; g.K = K; g.batch = batch;
  g.sAb = g.sAm = g.sAk = g.sBb = g.sBk = g.sBn = g.sCb = g.sCm = g.sCn = 0;
  g.lim = nullptr; g.lim_m = g.lim_n = g.lim_k = 0;
  g.alpha = 1.f; g.beta = 0.f; g.alpha_dev = nullptr; g.bias = nullptr; g.relu = 0; g.split_k = 0;
  return g;
}

static int pick_split(int M, int N, long long K) {
  const long long tiles = (long long)((M + 63) / 64) * ((N + 63) / 64);
  long long s = (3LL * kNumSMs + tiles - 1) / tiles;
  const long long maxs = K / 512 > 1 ? K / 512 : 1;
  if (s > maxs) s = maxs;
  if (s > 1024) s = 1024;
  return (int)(s < 1 ? 1 : s);
}

}  // namespace gp

using namespace gp;

extern "C" int gp_version(void) { return 100; }
extern "C" const char* gp_last_error(void) { return g_err; }
extern "C" long long gp_launch_count(void) { return g_launches.load(); }
extern "C" void gp_launch_count_reset(void) { g_launches.store(0); }

extern "C" int gp_graphconv_fwd(const float* x, long long ldx, const float* adj, const float* w, const float* bias,
                                const int32_t* nb, int B, int N, int din, int dout, int add_self, int normalize,
                                float* u, float* y, long long ldy, float* rnorm, int precision,
                                gp_stream_t stream) {
  GP_REQUIRE(x && adj && w && u && y, "graphconv_fwd: null pointer");
  GP_REQUIRE(B > 0 && N > 0 && din > 0 && dout > 0 && ldx >= din && ldy >= dout, "graphconv_fwd: bad dims");
  GP_REQUIRE(!normalize || rnorm, "graphconv_fwd: normalize needs rnorm");
  cudaStream_t st = S(stream);
  (void)precision;
  // ENZYMES-sized graphs: one CTA per graph runs the whole layer out of shared memory (small_gcn.cu)
  if (small_gcn_eligible(B, N, din, dout, add_self))
    return small_gcn_fwd(x, ldx, adj, w, bias, nb, B, N, din, dout, normalize, u, y, ldy, rnorm, st);
  // U = A.X (+X)
  if (add_self)
    GP_CUDA(cudaMemcpy2DAsync(u, (size_t)din * 4, x, (size_t)ldx * 4, (size_t)din * 4, (size_t)B * N,
                              cudaMemcpyDeviceToDevice, st));
  gp_gemm g = mk(adj, x, u, N, din, N, B);
  g.sAb = (long long)N * N; g.sAm = N; g.sAk = 1;
  g.sBb = (long long)N * ldx; g.sBk = ldx; g.sBn = 1;
  g.sCb = (long long)N * din; g.sCm = din; g.sCn = 1;
  g.lim = nb; g.lim_m = g.lim_k = nb != nullptr;
  g.beta = add_self ? 1.f : 0.f;
  GP_TRY(bgemm_f32(g, st));
  // V = U.W + b   (rows flattened over the batch)
  gp_gemm h = mk(u, w, y, B * N, dout, din, 1);
  h.sAm = din; h.sAk = 1; h.sBk = dout; h.sBn = 1; h.sCm = ldy; h.sCn = 1;
  h.bias = bias;
  GP_TRY(bgemm_f32(h, st));
  // Y = V / max(||V||, eps)
  if (normalize) GP_TRY(bias_normalize(y, nullptr, rnorm, (long long)B * N, dout, ldy, 1, st));
  return GP_OK;
}

extern "C" int gp_graphconv_bwd(const float* dv, const float* u, const float* x, long long ldx, const float* adj,
                                const float* w, const int32_t* nb, int B, int N, int din, int dout, int add_self,
                                float* dw, float* db, float* du, float* dx, float* dadj, float* ws,
                                int precision, gp_stream_t stream) {
  GP_REQUIRE(dv && u && w && dw, "graphconv_bwd: null pointer");
  GP_REQUIRE(B > 0 && N > 0 && din > 0 && dout > 0, "graphconv_bwd: bad dims");
  GP_REQUIRE((!dx && !dadj) || (du && adj), "graphconv_bwd: dx/dadj need du and adj");
  GP_REQUIRE(!dadj || x, "graphconv_bwd: dadj needs x");
  cudaStream_t st = S(stream);
  (void)precision;
  if (small_gcn_eligible(B, N, din, dout, add_self)) {       // ws: gp_graphconv_bwd_ws floats
    GP_REQUIRE(ws && x && adj, "graphconv_bwd: the per-graph fused path needs ws, x and adj");
    return small_gcn_bwd(dv, u, x, ldx, adj, w, nb, B, N, din, dout, dw, db, dx, dadj, ws, st);
  }
  const long long rows = (long long)B * N;
  if (db != nullptr) {
    GP_REQUIRE(ws, "graphconv_bwd: db needs ws");
    GP_TRY(colsum(dv, rows, dout, dout, db, 0, ws, st));
  }
  // dW = U^T dV  (reduction over all B*N rows, split-K)
  {
    GP_REQUIRE(rows <= 0x7fffffffLL, "graphconv_bwd: B*N too large");
    gp_gemm g = mk(u, dv, dw, din, dout, (int)rows, 1);
    g.sAm = 1; g.sAk = din; g.sBk = dout; g.sBn = 1; g.sCm = dout; g.sCn = 1;
    g.split_k = pick_split(din, dout, rows);
    GP_TRY(bgemm_f32(g, st));
  }
  if (dx == nullptr && dadj == nullptr) return GP_OK;
  // dU = dV W^T
  {
    gp_gemm g = mk(dv, w, du, (int)rows, din, dout, 1);
    g.sAm = dout; g.sAk = 1; g.sBk = 1; g.sBn = dout; g.sCm = din; g.sCn = 1;
    GP_TRY(bgemm_f32(g, st));
  }
  if (dx != nullptr) {      // dX = A^T dU (+dU)
    if (add_self) GP_CUDA(cudaMemcpyAsync(dx, du, (size_t)rows * din * 4, cudaMemcpyDeviceToDevice, st));
    gp_gemm g = mk(adj, du, dx, N, din, N, B);
    g.sAb = (long long)N * N; g.sAm = 1; g.sAk = N;
    g.sBb = (long long)N * din; g.sBk = din; g.sBn = 1;
    g.sCb = (long long)N * din; g.sCm = din; g.sCn = 1;
    g.lim = nb; g.lim_m = g.lim_k = nb != nullptr;
    g.beta = add_self ? 1.f : 0.f;
    GP_TRY(bgemm_f32(g, st));
  }
  if (dadj != nullptr) {    // dA += dU X^T
    gp_gemm g = mk(du, x, dadj, N, N, din, B);
    g.sAb = (long long)N * din; g.sAm = din; g.sAk = 1;
    g.sBb = (long long)N * ldx; g.sBk = 1; g.sBn = ldx;
    g.sCb = (long long)N * N; g.sCm = N; g.sCn = 1;
    g.beta = 1.f;
    GP_TRY(bgemm_f32(g, st));
  }
  return GP_OK;
}

extern "C" int gp_pool_fwd(const float* s, const float* z, long long ldz, const float* adj, const int32_t* nb,
                           int B, int N, int K, int F, float* xp, float* t, float* ap, int precision,
                           gp_stream_t stream) {
  GP_REQUIRE(s && z && adj && xp && t && ap, "pool_fwd: null pointer");
  GP_REQUIRE(B > 0 && N > 0 && K > 0 && F > 0 && ldz >= F, "pool_fwd: bad dims");
  cudaStream_t st = S(stream);
  (void)precision;
  // ENZYMES-sized graphs: one CTA per graph chains S^T Z, S^T A and (S^T A) S out of shared memory (small_gcn.cu)
  if (small_pool_eligible(B, N, K, F)) return small_pool_fwd(s, z, ldz, adj, nb, B, N, K, F, xp, t, ap, st);
  const int lim = nb != nullptr;
  {   // X' = S^T Z
    gp_gemm g = mk(s, z, xp, K, F, N, B);
    g.sAb = (long long)N * K; g.sAm = 1; g.sAk = K;
    g.sBb = (long long)N * ldz; g.sBk = ldz; g.sBn = 1;
    g.sCb = (long long)K * F; g.sCm = F; g.sCn = 1;
    g.lim = nb; g.lim_k = lim;
    GP_TRY(bgemm_f32(g, st));
  }
  {   // T = S^T A
    gp_gemm g = mk(s, adj, t, K, N, N, B);
    g.sAb = (long long)N * K; g.sAm = 1; g.sAk = K;
    g.sBb = (long long)N * N; g.sBk = N; g.sBn = 1;
    g.sCb = (long long)K * N; g.sCm = N; g.sCn = 1;
    g.lim = nb; g.lim_k = lim; g.lim_n = lim;
    GP_TRY(bgemm_f32(g, st));
  }
  {   // A' = T S
    gp_gemm g = mk(t, s, ap, K, K, N, B);
    g.sAb = (long long)K * N; g.sAm = N; g.sAk = 1;
    g.sBb = (long long)N * K; g.sBk = K; g.sBn = 1;
    g.sCb = (long long)K * K; g.sCm = K; g.sCn = 1;
    g.lim = nb; g.lim_k = lim;
    GP_TRY(bgemm_f32(g, st));
  }
  return GP_OK;
}

extern "C" int gp_pool_bwd(const float* dxp, const float* dap, const float* s, const float* z, long long ldz,
                           const float* adj, const float* t, const int32_t* nb, int B, int N, int K, int F,
                           float* dz, long long lddz, int accumulate_dz, float* ds, int accumulate_ds,
                           float* dadj, float* ws, int precision, gp_stream_t stream) {
  GP_REQUIRE(dxp && dap && s && z && adj && t && dz && ds && ws, "pool_bwd: null pointer");
  GP_REQUIRE(B > 0 && N > 0 && K > 0 && F > 0 && ldz >= F && lddz >= F, "pool_bwd: bad dims");
  cudaStream_t st = S(stream);
  (void)precision;
  if (dadj == nullptr && small_pool_eligible(B, N, K, F))
    return small_pool_bwd(dxp, dap, s, z, ldz, adj, t, nb, B, N, K, F, dz, lddz, accumulate_dz, ds, accumulate_ds, st);
  const int lim = nb != nullptr;
  {   // dZ (+)= S dX'
    gp_gemm g = mk(s, dxp, dz, N, F, K, B);
    g.sAb = (long long)N * K; g.sAm = K; g.sAk = 1;
    g.sBb = (long long)K * F; g.sBk = F; g.sBn = 1;
    g.sCb = (long long)N * lddz; g.sCm = lddz; g.sCn = 1;
    g.lim = nb; g.lim_m = lim;
    g.beta = accumulate_dz ? 1.f : 0.f;
    GP_TRY(bgemm_f32(g, st));
  }
  {   // dS (+)= Z dX'^T
    gp_gemm g = mk(z, dxp, ds, N, K, F, B);
    g.sAb = (long long)N * ldz; g.sAm = ldz; g.sAk = 1;
    g.sBb = (long long)K * F; g.sBk = 1; g.sBn = F;
    g.sCb = (long long)N * K; g.sCm = K; g.sCn = 1;
    g.lim = nb; g.lim_m = lim;
    g.beta = accumulate_ds ? 1.f : 0.f;
    GP_TRY(bgemm_f32(g, st));
  }
  {   // dS += T^T dA'
    gp_gemm g = mk(t, dap, ds, N, K, K, B);
    g.sAb = (long long)K * N; g.sAm = 1; g.sAk = N;
    g.sBb = (long long)K * K; g.sBk = K; g.sBn = 1;
    g.sCb = (long long)N * K; g.sCm = K; g.sCn = 1;
    g.lim = nb; g.lim_m = lim;
    g.beta = 1.f;
    GP_TRY(bgemm_f32(g, st));
  }
  {   // ws = S dA'^T ;  dS += A ws
    gp_gemm g = mk(s, dap, ws, N, K, K, B);
    g.sAb = (long long)N * K; g.sAm = K; g.sAk = 1;
    g.sBb = (long long)K * K; g.sBk = 1; g.sBn = K;
    g.sCb = (long long)N * K; g.sCm = K; g.sCn = 1;
    g.lim = nb; g.lim_m = lim;
    GP_TRY(bgemm_f32(g, st));
    gp_gemm h = mk(adj, ws, ds, N, K, N, B);
    h.sAb = (long long)N * N; h.sAm = N; h.sAk = 1;
    h.sBb = (long long)N * K; h.sBk = K; h.sBn = 1;
    h.sCb = (long long)N * K; h.sCm = K; h.sCn = 1;
    h.lim = nb; h.lim_m = lim; h.lim_k = lim;
    h.beta = 1.f;
    GP_TRY(bgemm_f32(h, st));
  }
  if (dadj != nullptr) {   // dA += S dA' S^T   (pooled levels whose input adjacency needs a gradient)
    gp_gemm g = mk(s, dap, ws, N, K, K, B);
    g.sAb = (long long)N * K; g.sAm = K; g.sAk = 1;
    g.sBb = (long long)K * K; g.sBk = K; g.sBn = 1;
    g.sCb = (long long)N * K; g.sCm = K; g.sCn = 1;
    GP_TRY(bgemm_f32(g, st));
    gp_gemm h = mk(ws, s, dadj, N, N, K, B);
    h.sAb = (long long)N * K; h.sAm = K; h.sAk = 1;
    h.sBb = (long long)N * K; h.sBk = 1; h.sBn = K;
    h.sCb = (long long)N * N; h.sCm = N; h.sCn = 1;
    h.beta = 1.f;
    GP_TRY(bgemm_f32(h, st));
  }
  return GP_OK;
}
