/* gp_b200.h -- C ABI of the B200-native DiffPool hot path (libgp_b200.so).
 *
 * Drop-in boundary for the hot path of JiaxuanYou/graph-pooling (encoders.py:976-1334 plus the
 * DiffPool GraphConv at encoders.py:296-328).  The reference has no native code and no FFI; the
 * functions below are what a binding for this path would bind, one per reference function or per
 * fused group of ATen calls on the path.  Every pointer is a DEVICE pointer unless the name ends
 * in `_host`; tensors are row-major fp32; `nb` is the per-graph node count (`batch_num_nodes`,
 * train.py:200) as int32 on the device, or NULL for "no masking / no tile skipping".
 *
 * Conventions
 *  - every function returns 0 on success and a negative gp_status on failure, never throws;
 *    gp_last_error() returns a thread-local message for the last failure;
 *  - every function is asynchronous on `stream` (a cudaStream_t passed as void*);
 *  - no function allocates device memory: workspaces are passed in (sizes documented per call);
 *  - no global mutable state except a lazily initialised, read-only device-attribute cache.
 *
 * Zero-padding contract (graph_sampler.py:97-109): when `nb` is given, rows and columns of the
 * level-0 adjacency at index >= nb[b] are zero.  The N^2-sized contractions use this to skip
 * all-zero tiles; every other op processes pad rows densely, exactly as the reference does
 * (pad rows carry normalize(bias) -> ReLU -> BN values and take part in the BN statistics,
 * encoders.py:1061-1080).
 */
#ifndef GP_B200_H
#define GP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* gp_stream_t; /* cudaStream_t */

enum gp_status {
  GP_OK = 0,
  GP_ERR_INVALID = -1,   /* bad argument (shape, stride, null pointer)      */
  GP_ERR_CUDA = -2,      /* a CUDA runtime call or launch failed             */
  GP_ERR_UNSUPPORTED = -3 /* valid request outside what the kernel supports */
};

/* precision of the dense contractions */
enum gp_precision {
  GP_F32 = 0,    /* fp32 FFMA everywhere (parity anchor, <=1e-5 relative)                       */
  GP_BF16 = 1    /* bf16 operands on tcgen05 tensor cores, fp32 accumulation in TMEM             */
  /* (a split-operand mode -- {0,1} adjacency exact in bf16, real operands as hi+lo pairs accumulated through the
   * multi-pair GEMM for ~fp32 accuracy on tensor cores -- is NOT implemented in this round) */
};

int gp_version(void);
const char* gp_last_error(void);
/* number of kernels launched by this library (all threads of the process: autograd runs the backward on
 * its own thread) since the last reset (bench.py's `gpu_launches`). */
long long gp_launch_count(void);
void gp_launch_count_reset(void);

/* ---------------------------------------------------------------------------------------------
 * Generic strided batched GEMM  C[b] = alpha * op(A[b]) . op(B[b]) (+bias) (relu) + beta*C[b]
 * Replaces every aten::bmm / aten::mm / addmm on the path (torch.matmul at encoders.py:319,322,
 * 1278,1279,1311; nn.Linear at :1024-1031).  Transposes are expressed through strides.
 * With `lim` != NULL, lim_m/lim_n/lim_k select which logical dims are clipped to lim[b] for
 * batch b (tile skipping); outputs outside the clipped range are written as beta*C (+0).
 * split_k > 1 accumulates partial sums with atomics into C, which must hold the beta-term
 * already (the wrapper zero-fills it when beta == 0).
 * ------------------------------------------------------------------------------------------- */
typedef struct gp_gemm {
  const float* A; const float* B; float* C;
  int M, N, K, batch;
  long long sAb, sAm, sAk;   /* element strides of A: batch, m, k */
  long long sBb, sBk, sBn;
  long long sCb, sCm, sCn;
  const int32_t* lim;        /* per-batch limit or NULL */
  int lim_m, lim_n, lim_k;   /* 1 = clip that dim to lim[b] */
  float alpha, beta;
  const float* alpha_dev;    /* optional device scalar multiplied into alpha */
  const float* bias;         /* optional [N] added before relu */
  int relu;
  int split_k;               /* 0/1 = none */
} gp_gemm;
int gp_bgemm_f32(const gp_gemm* g, gp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Tensor-core variant of the same contraction: bf16 operands (TMA -> shared memory, 128B swizzle),
 * tcgen05.mma with fp32 accumulation in TMEM, fp32 output C and/or a bf16 copy Cb (so the next
 * contraction needs no conversion pass).  Operands keep their natural row-major layout:
 *   a_major 0: A stored [M rows][K cols] (K-major)   1: A stored [K rows][M cols] (M-major, "A^T")
 *   b_major 0: B stored [N rows][K cols] (K-major)   1: B stored [K rows][N cols] (N-major)
 * ldA/ldB/sAb/sBb in elements, multiples of 8; bases 16-byte aligned (TMA).  lim_k skips whole
 * 64-wide K tiles, it does not mask inside a tile: between lim[b] and the next multiple of 64 one
 * operand must be ZERO and the other finite (true on the DiffPool path: zero-padded adjacency,
 * masked S), otherwise use gp_bgemm_f32.
 * ------------------------------------------------------------------------------------------- */
typedef struct gp_gemm_bf16 {
  const void* A; const void* B;          /* bf16 */
  float* C; void* Cb;                    /* fp32 and/or bf16 output (either may be NULL) */
  int M, N, K, batch;
  long long ldA, sAb; int a_major;
  long long ldB, sBb; int b_major;
  long long ldC, sCb, ldCb, sCbb;
  const int32_t* lim; int lim_m, lim_n, lim_k;
  float alpha, beta; const float* alpha_dev;
  const float* bias; int relu;
  int split_k;
} gp_gemm_bf16;
int gp_bgemm_bf16(const gp_gemm_bf16* g, gp_stream_t stream);

/* Persistent multi-pair form (the one the engine uses):
 *   C[b] = alpha * sum_{q < npairs} op(A_q[b]) . op(B_q[b]) (+bias)(relu) + beta * C[b]
 * One launch accumulates up to 4 different products into the same TMEM accumulator tile, e.g. the
 * pooling backward dS = Z dX'^T + T^T dA' + A (S dA'^T) (encoders.py:1278-1279) or the link-loss
 * backward dS = G S + G^T S, instead of read-modify-write chains over [B,N,K] buffers.  CTAs are
 * persistent (one per SM), the accumulator is double-buffered in TMEM so one tile's epilogue overlaps
 * the next tile's MMAs, and the epilogue stores are staged through shared memory (coalesced).
 * Same operand rules as gp_gemm_bf16; lim_k is per pair. */
typedef struct gp_operand_pair {
  const void* A; const void* B;          /* bf16 */
  int K;
  long long ldA, sAb; int a_major;
  long long ldB, sBb; int b_major;
  int lim_k;
} gp_operand_pair;
typedef struct gp_gemm_bf16x {
  gp_operand_pair pair[4]; int npairs;
  float* C; void* Cb;
  int M, N, batch;
  long long ldC, sCb, ldCb, sCbb;
  const int32_t* lim; int lim_m, lim_n;
  float alpha, beta; const float* alpha_dev;
  const float* bias; int relu;
  int split_k;
  /* device-side switch, evaluated by the kernel (no host sync): if cond != NULL and *cond == 0, only the first
   * cond_npairs pairs are accumulated and alpha is multiplied by cond_alpha; cond_npairs == 0 makes the launch a
   * no-op (C untouched).  The DiffPool backward passes gp_adj_prepare's "adjacency is not symmetric" flag: for a
   * symmetric A, (G + G^T).S = 2 G.S and T^T dA' + A.(S dA'^T) = T^T (dA' + dA'^T). */
  const int32_t* cond; int cond_npairs; float cond_alpha;
  /* optional device permutation of the batch index (e.g. argsort(-nb)): ragged batches are walked from the largest
   * graph to the smallest, so the static round-robin over persistent CTAs stays balanced (longest-first). */
  const int32_t* order;
  /* tri = 1 (two pairs reading the SAME symmetric M x M operand, e.g. dS = (G + G^T) S with the link-loss gradient G;
   * needs cond with cond_npairs == 2): when *cond == 0 the operand is taken to be stored as its upper diagonal band
   * only (what gp_linkloss_tc writes in mode 2): for the row block starting at m0, pair 0 (A K-major) contracts over
   * k >= c0 and pair 1 (the same buffer read M-major, i.e. transposed) over k < c0, with c0 = floor(m0 / 256) * 256 --
   * together one full contraction, so cond_alpha carries the factor 2.  When *cond != 0 both pairs run in full. */
  int tri;
} gp_gemm_bf16x;
int gp_bgemm_bf16x(const gp_gemm_bf16x* g, gp_stream_t stream);
/* Fused GraphConv tail on tensor cores (encoders.py:322-326): one operand pair, batch == 1, N <= 256:
 *   V = alpha * A.B + bias ;  Y = V / max(||V||_2, 1e-12) per row  -> C (fp32) and/or Cb (bf16)
 *   rnorm[M] (optional) = the divisor;  rowstat[M][2] (optional) = (sum_c r, sum_c r^2) with r = relu(Y) if
 *   stat_relu else Y: the per-row sums from which gp_bn_finalize derives the BatchNorm-per-node statistics
 *   (encoders.py:1062-1064) without another pass over Y. */
int gp_bgemm_bf16_norm(const gp_gemm_bf16x* g, float* rnorm, float* rowstat, int stat_relu, gp_stream_t stream);
/* mean[n], invstd[n] over (batch, feature) from rowstat[b*N + n] (biased variance, eps 1e-5). */
int gp_bn_finalize(const float* rowstat, int B, int N, int d, float* mean, float* invstd, gp_stream_t stream);
/* H[b,n,:] = (relu?(Y[b,n,:]) - mean[n]) * invstd[n] -> fp32 h (row stride ldh) and/or bf16 copy. */
/* h_bf16_2 (optional): a second bf16 destination -- the column half of the operand that feeds ONE A.X
 * contraction for two GCN stacks sharing the adjacency (embedding + assignment GCN). */
int gp_bn_apply(const float* y, long long ldy, const float* mean, const float* invstd, int B, int N, int d,
                int relu, int bn, float* h, long long ldh, void* h_bf16, long long ldhb, void* h_bf16_2,
                long long ldhb2, gp_stream_t stream);
/* gp_bias_normalize_f32 / gp_softmax_mask_fwd / _bwd with the bf16 operand copy written in the same pass;
 * softmax backward can also return dcol = colsum(dT) (the assign_pred bias gradient, encoders.py:1273);
 * ws >= (148*16 + 256) * K floats.  The bf16 copies need 16-byte aligned rows, width % 4 == 0 and <= 2048. */
int gp_bias_normalize_x(float* v, const float* bias, float* rnorm, long long rows, int d, long long ld,
                        int normalize, void* y_bf16, long long ldyb, gp_stream_t stream);
int gp_softmax_mask_fwd_x(float* t, const int32_t* nb, int B, int N, int K, void* s_bf16, long long ldsb,
                          gp_stream_t stream);
int gp_softmax_mask_bwd_x(const float* s, const float* ds, const int32_t* nb, int B, int N, int K, float* dt,
                          void* dt_bf16, long long lddtb, float* dcol, float* ws, gp_stream_t stream);
/* Adjacency preparation (one HBM pass): adj [B,N,N] fp32 (adj_dtype 0, train.py:197) or uint8 {0,1} (adj_dtype 1,
 * compact feed) -> bf16 operand [B,N,ld] (N <= ld < N+32, zero padded); flags[0] != 0 iff some graph's adjacency
 * is NOT symmetric, flags[1] != 0 iff some entry is outside {0,1}.  With nb, tiles beyond nb[b] are written as
 * zeros without being read (feed contract graph_sampler.py:97-109). */
int gp_adj_prepare(const void* adj, int adj_dtype, const int32_t* nb, int B, int N, void* adj_bf16, long long ld,
                   int32_t* flags, gp_stream_t stream);
/* Edge-list feed: the bf16 operand [B,N,ld] straight from the batch's edge lists -- edges [E][2]
 * graph-local node ids of id_bytes = 4 (int32) or 2 (uint16, N <= 65536) bytes each, graph b owns edges
 * eptr[b] .. eptr[b+1]; undirected = 1 writes both (u,v) and (v,u).  flags as gp_adj_prepare: [0] = 0 for an
 * undirected list (symmetric by construction), [1] = 0.  What graph_sampler.py:97-109 + train.py:197-201 ship as
 * B*N*N*4 bytes becomes 4 or 8 bytes per edge. */
int gp_adj_from_edges(const void* edges, int id_bytes, const int32_t* eptr, int B, int N, int max_edges_per_graph,
                      int undirected, void* adj_bf16, long long ld, int32_t* flags, int accumulate_flags,
                      gp_stream_t stream);
/* Extended form: adj_dtype 2 = bit-packed rows from gp_host_pack_adj_bits (bit c & 7 of byte c >> 3); ld_in = input
 * row stride in elements (bytes for dtype 2; 0 = dense); accumulate_flags = 1 ORs into `flags` instead of
 * resetting them, so a batch may be prepared in several calls (e.g. one part fed as bits, one as fp32). */
int gp_adj_prepare_x(const void* adj, int adj_dtype, long long ld_in, const int32_t* nb, int B, int N,
                     void* adj_bf16, long long ld, int32_t* flags, int accumulate_flags, gp_stream_t stream);
/* HOST function (no CUDA): packs `rows` rows of N fp32 entries (a dense {0,1} adjacency in host memory, the
 * reference's feed, train.py:197) into bits, `ldb` bytes per row, on `threads` host threads (AVX2 when available).
 * *non01 = 1 if any entry is outside {0,1} (the caller must then feed those rows as fp32).  Shrinks the PCIe
 * transfer 32x; 16 threads pack ~87 GB/s on the B200 box (PCIe moves the raw floats at ~65 GB/s). */
int gp_host_pack_adj_bits(const float* adj_host, long long rows, int N, void* out_host, long long ldb, int threads,
                          int* non01);
/* out[b] = bf16((cond && *cond == 0) ? x[b] + x[b]^T : x[b]), x [B,K,K] fp32 with row stride ldx, out row stride
 * ld (zero padded). */
int gp_sym_select_bf16(const float* x, long long ldx, int B, int K, const int32_t* cond, void* out_bf16,
                       long long ld, gp_stream_t stream);
/* y[r, 0:cols_pad] = bf16(x[r, 0:cols]) zero-padded to cols_pad (row strides ldx / ldy in elements) */
int gp_cvt_f32_bf16(const float* x, long long ldx, void* y, long long ldy, long long rows, int cols,
                    int cols_pad, gp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * GraphConv  (encoders.py:315-328):  U = A.X (+X);  V = U.W + b;  Y = V / max(||V||_2, 1e-12)
 *   x [B,N,din] row stride ldx;  adj [B,N,N];  w [din,dout];  bias [dout] or NULL
 *   u [B,N,din] (saved for backward);  y [B,N,dout] row stride ldy;  rnorm [B,N] = max(||V||,eps)
 * ------------------------------------------------------------------------------------------- */
int gp_graphconv_fwd(const float* x, long long ldx, const float* adj, const float* w, const float* bias,
                     const int32_t* nb, int B, int N, int din, int dout, int add_self, int normalize,
                     float* u, float* y, long long ldy, float* rnorm, int precision, gp_stream_t stream);

/* backward of GraphConv given dV (gradient w.r.t. the pre-normalisation V, see gp_gcn_layer_bwd):
 *   dW = sum_b U^T dV ; db = colsum dV ; dU = dV W^T ; dX = A^T dU (+dU) ; dA += dU X^T
 *   dx / dadj may be NULL (not needed).  du: workspace [B,N,din].  ws: gp_graphconv_bwd_ws(...) floats.
 * ENZYMES-sized shapes (N, din, dout <= 128, add_self == 0, operands fitting one SM's shared memory) run as ONE
 * CTA per graph that stages the real n_b x n_b adjacency block, X, U, W once and computes the whole layer
 * (forward) / all four gradient products (backward; per-graph dW / db partials reduced deterministically). */
long long gp_graphconv_bwd_ws(int B, int N, int din, int dout, int add_self);
int gp_graphconv_bwd(const float* dv, const float* u, const float* x, long long ldx, const float* adj,
                     const float* w, const int32_t* nb, int B, int N, int din, int dout, int add_self,
                     float* dw, float* db, float* du, float* dx, float* dadj, float* ws,
                     int precision, gp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * ReLU + BatchNorm1d(num_nodes) with batch statistics (encoders.py:1062-1064,1048-1052):
 *   H[b,n,:] = (relu(Y[b,n,:]) - mean_n) * invstd_n, statistics over (b, feature) per node index.
 *   y [B,N,d] contiguous; h row stride ldh (a column slot of the concat buffer, :1078).
 *   relu/bn flags switch the two stages (last layer: neither -> plain copy).
 * ------------------------------------------------------------------------------------------- */
int gp_relu_bn_fwd(const float* y, float* h, long long ldh, float* mean, float* invstd,
                   int B, int N, int d, int relu, int bn, gp_stream_t stream);
/* Same, with a workspace of gp_relu_bn_fwd_ws(B, N, d) floats (may be 0): for SMALL N and LARGE B (ENZYMES-sized
 * graphs in big batches) one CTA per node index would leave most SMs idle, so each node's batch is cut into
 * slices (grid N x S) whose statistics are combined exactly (Chan's parallel mean / M2 update). */
long long gp_relu_bn_fwd_ws(int B, int N, int d);
int gp_relu_bn_fwd_x(const float* y, float* h, long long ldh, float* mean, float* invstd,
                     int B, int N, int d, int relu, int bn, float* ws, gp_stream_t stream);

/* Backward through [concat slot + readout scatter + next-layer dX] -> BN -> ReLU -> normalize:
 *   g  = dz (dense slot gradient, row stride lddz, or NULL)
 *      + dxn (gradient from the next layer's A^T dU, contiguous [B,N,d], or NULL)
 *      + scatter(dout[b, c] at row argidx[b, c]) (max-readout backward, or NULL)
 *   dR = (g - mean(g) - Hhat*mean(g*Hhat)) * invstd   (if bn);  dY = dR*[Y>0] (if relu)
 *   dV = (dY - Y<Y,dY>)/r  (if normalize; rows whose norm was clamped: dV = dY/eps)
 *   h: the forward output slot (Hhat), row stride ldh; y: forward Y (contiguous; for the last
 *   layer y == h slot with ldy == ldh); dv [B,N,d] contiguous. */
int gp_gcn_layer_bwd(const float* dz, long long lddz, const float* dxn, const float* dout,
                     const int32_t* argidx, long long ldo, const float* h, long long ldh,
                     const float* y, long long ldy, const float* rnorm, const float* invstd,
                     int B, int N, int d, int relu, int bn, int normalize, float* dv,
                     gp_stream_t stream);

/* Extended form (the one the engines use): optional bf16 copy of dV (the operand of the tensor-core dW / dU
 * contractions), the bias gradient db = colsum(dV) produced in the same pass, and Hhat recomputed from Y and
 * the saved statistics (h == NULL, mean != NULL) so the BN output need not be re-read.  16-byte aligned
 * inputs with d in {32, 64, 128} (bn) or d % 4 == 0, d <= 2048 (no bn; two passes over a row beyond 512) take a vectorised kernel
 * (a thread-block cluster per node index, batch means reduced through distributed shared memory).
 * ws: gp_gcn_layer_bwd_ws_x(q) floats, needed when db != NULL (gp_gcn_layer_bwd_ws(B, N, d, bn) is the
 * shape-only bound, valid when every operand meets the alignment above or dv != NULL); optional otherwise --
 * with it the generic kernel may split a node's batch over several CTAs (small N, large B). */
typedef struct gp_layer_bwd {
  const float* dz; long long lddz;
  const float* dxn;
  const float* dout; const int32_t* argidx; long long ldo;
  const float* h; long long ldh;
  const float* y; long long ldy;
  const float* rnorm; const float* mean; const float* invstd;
  int B, N, d, relu, bn, normalize;
  float* dv; void* dv_bf16; long long lddvb;
  float* db; float* ws;
  long long lddxn;           /* row stride of dxn in elements; 0 = contiguous (d) */
  const int32_t* nb_zero;    /* optional, layers without BN only: the caller guarantees that rows n >= nb_zero[b] receive
                                no upstream gradient (masked level: encoders.py:1080, 1275) -- their dV is written as
                                zero without reading the operands (padding-aware row pass) */
  int dz_bf16, dxn_bf16;     /* 1: dz / dxn point to bf16 data (row strides still in elements): the tensor-core schedule
                                keeps these gradient intermediates (dza, dZ, the lock-step dX) in bf16 -- half the bytes
                                to write and to read; vectorised path only */
} gp_layer_bwd;
int gp_gcn_layer_bwd_x(const gp_layer_bwd* q, gp_stream_t stream);
long long gp_gcn_layer_bwd_ws(int B, int N, int d, int bn);
long long gp_gcn_layer_bwd_ws_x(const gp_layer_bwd* q);
/* 1 if a layer of this shape runs on the vectorised kernels (given 16-byte aligned operands) -- the precondition of
 * dz_bf16 / dxn_bf16. */
int gp_gcn_layer_bwd_vectorised(int B, int d, int bn);
/* 1 if this shape is served by a kernel instantiated for bf16 gradient sources (otherwise they are slower than fp32). */
int gp_gcn_layer_bwd_bf16_sources_fast(int B, int d, int bn);

/* ---------------------------------------------------------------------------------------------
 * Max readout (encoders.py:1097,1257,1287): out[b,f] = max_n Z[b,n,f], pad rows (n >= nb[b])
 * counting as 0 when nb != NULL (the mask of :1078-1080).  argidx = winning row (lowest index on
 * ties, as torch CPU), or -1 if a pad row won (its gradient is dropped by the mask).
 * ------------------------------------------------------------------------------------------- */
int gp_readout_max_fwd(const float* z, long long ldz, const int32_t* nb, int B, int N, int F,
                       float* out, int32_t* argidx, long long ldo, gp_stream_t stream);
/* Same over the first N rows of graphs stored `pitch` rows apart (pitch >= N): rows [N, pitch) are not part of the
 * readout at all.  Used for pooled levels whose cluster count was padded to a multiple of 8 (dead clusters). */
int gp_readout_max_fwd_x(const float* z, long long ldz, int pitch, const int32_t* nb, int B, int N, int F,
                         float* out, int32_t* argidx, long long ldo, gp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Assignment softmax (encoders.py:1273-1275): in place, S = softmax(T) on rows n < nb[b], 0 on
 * pad rows.  Backward: dT = S*(dS - <dS,S>) on real rows, 0 on pad rows.
 * ------------------------------------------------------------------------------------------- */
int gp_softmax_mask_fwd(float* t, const int32_t* nb, int B, int N, int K, gp_stream_t stream);
int gp_softmax_mask_bwd(const float* s, const float* ds, const int32_t* nb, int B, int N, int K,
                        float* dt, gp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Chained pooling contraction on tensor cores (encoders.py:1279): A' = S^T A S in ONE launch, the
 * intermediate T = S^T A staying ON CHIP: a T tile [128 clusters x 128 nodes] is accumulated in TMEM,
 * converted to bf16 into shared memory in UMMA operand layout and is the A operand of a second
 * tcgen05.mma that accumulates the A' row block [128 x K] in another TMEM region.  K <= 256: one CTA per
 * row block.  256 < K <= 512 (the A' row block alone would fill the SM's 512 TMEM columns): a cluster of two
 * CTAs shares a row block, each accumulates one column half of A', computes every second T tile and pushes
 * it into the partner's shared memory (bulk DSMEM copy), so every T tile is still computed once.
 *   s   [B,N,lds] bf16 (masked assignment, pad rows zero), adj [B,N,ldadj] bf16, nb (optional) node counts,
 *   order (optional) batch permutation for ragged batches (longest first).
 *   t   (optional, bf16 [B,K,ldt]): training keeps T for the backward (dS += T^T dA'); it is written once
 *       from the shared-memory tile by a TMA store.  NULL (inference): T never exists in HBM.
 *       Columns beyond ceil(nb/128)*128 are left untouched (never read: every consumer clips to nb).
 *   ap  (optional fp32 [B,K,ldap]) / ap_bf16 (optional [B,K,ldapb]): A'.
 * Strides in elements, operand strides multiples of 8, bases 16-byte aligned.  K <= 512.
 * ------------------------------------------------------------------------------------------- */
int gp_pool_chain_bf16(const void* s_bf16, long long lds, const void* adj_bf16, long long ldadj,
                       const int32_t* nb, const int32_t* order, int B, int N, int K, void* t_bf16,
                       long long ldt, float* ap, long long ldap, void* ap_bf16, long long ldapb,
                       gp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Pooling (encoders.py:1278-1279):  X' = S^T Z ;  T = S^T A ;  A' = T S
 *   z row stride ldz (concat buffer, pad rows need no masking because S's pad rows are 0).
 *   t [B,K,N] is saved for backward.
 * Backward:  dZ (+)= S dX' ;  dS (+)= Z dX'^T + T^T dA' + A (S dA'^T) ;  dA = S dA' S^T (if dadj)
 *   ws: workspace >= B*N*K floats.  accumulate_ds: 1 = dS already holds the link-loss term.
 * ------------------------------------------------------------------------------------------- */
int gp_pool_fwd(const float* s, const float* z, long long ldz, const float* adj, const int32_t* nb,
                int B, int N, int K, int F, float* xp, float* t, float* ap, int precision,
                gp_stream_t stream);
int gp_pool_bwd(const float* dxp, const float* dap, const float* s, const float* z, long long ldz,
                const float* adj, const float* t, const int32_t* nb, int B, int N, int K, int F,
                float* dz, long long lddz, int accumulate_dz, float* ds, int accumulate_ds,
                float* dadj, float* ws, int precision, gp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Link-prediction loss (encoders.py:1311-1331), masked BCE between P = S S^T and A:
 *   partial[] <- per-tile sums of  -A log(P+1e-7) - (1-A) log(1-P+1e-7)  over the nb x nb block
 *   gsym (optional, [B,N,N]) <- dl/dP + (dl/dP)^T, un-normalised, for the backward
 *   n_partial = B * ceil(N/64)^2 floats.
 * Backward: dS = (upstream * inv_entries) * gsym . S    (gp_bgemm_f32 with alpha_dev).
 * gp_loss_finalize: total = ce (device scalar or NULL) + sum(partial) * inv_entries; link likewise.
 * ------------------------------------------------------------------------------------------- */
int gp_linkloss_fwd(const float* s, const float* adj, const int32_t* nb, int B, int N, int K,
                    float* partial, float* gsym, gp_stream_t stream);
int gp_loss_finalize(const float* partial, int n_partial, double inv_entries, const float* ce,
                     float* total, float* link, gp_stream_t stream);
/* Fused tensor-core form: S (bf16, row stride lds) -> P = S S^T tiles in TMEM -> masked BCE against
 * the bf16 adjacency in the epilogue; G = dl/dP evaluated with a[m,n] (bf16, row stride ldg, may be
 * NULL; the backward is dS = (G + G^T) S = two operand pairs of gp_bgemm_bf16x) and one partial per
 * epilogue warp: n_partial = gp_linkloss_tc_partials(B, N).  P never touches HBM.  s / adj / gsym
 * bases 16-byte aligned with row strides that are multiples of 8 take the 16-byte vector path.
 * gp_loss_finalize with n_partial > 8192 uses 256 floats of scratch AFTER the partial array. */
int gp_linkloss_tc_partials(int B, int N);
/* mode 0: masked BCE (the reference).  mode 1: Frobenius option -- partial sums of (A - P)^2 (per graph
 * contiguous: gp_linkloss_tc_partials(1, N) each, feed gp_frob_finalize) and G = -(A - P). */
/* adj_flags (optional): gp_adj_prepare's flags[2].  flags[1] == 0 (entries in {0,1}) selects the one-log / one-rcp
 * form for the whole launch (without flags the test is made per 32x32 chunk).  flags[0] == 0 (every adjacency
 * symmetric, BCE mode): P and G are symmetric, so only the tiles of the upper diagonal band are computed (56 % of
 * them at N = 2048, enumerated compactly so the persistent CTAs stay balanced); their epilogues also write the
 * transposed G chunks of the skipped tiles and count those loss terms twice. */
int gp_linkloss_tc(const void* s_bf16, long long lds, const void* adj_bf16, long long ldadj,
                   const int32_t* nb, int B, int N, int K, float* partial, void* gsym_bf16, long long ldg,
                   int mode, const int32_t* adj_flags, gp_stream_t stream);
/* adj_hop > 1 (encoders.py:1312-1317): Q = sum_{h=1..hop} (S S^T)^h [B,N,N] fp32 is formed by the caller's GEMM chain;
 * this pass clamps (min(Q,1), R3), reduces the masked BCE against fp32 `adj` into gp_linkloss_from_q_partials(B,N)
 * partial sums and writes G = dl/dQ (fp32 [B,N,N], zero where clamped or outside the n_b x n_b block; may be NULL). */
int gp_linkloss_from_q_partials(int B, int N);
int gp_linkloss_from_q(const float* Q, const float* adj, const int32_t* nb, int B, int N, float* partial, float* G,
                       gp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * North-star loss options that the reference does NOT contain (oracle: the DiffPool paper's definitions,
 * oracle/diffpool_oracle.py frobenius_link_loss / row_entropy_loss; parity unpinned by the reference).
 *
 * Frobenius link loss  L_F = mean_b || (A_b - S_b S_b^T) over the n_b x n_b block ||_F :
 *   gp_frob_link_fwd   partial[B * ceil(N/64)^2] <- per-tile sums of d^2 (per graph contiguous);
 *                      gsym (optional, [B,N,N]) <- -(d + d^T), un-normalised
 *   gp_frob_finalize   norm[b] = sqrt(sum of graph b's `per_graph` partials); coef[b] = 1/(B*norm[b]);
 *                      link = mean_b norm[b]; total = ce (device scalar or NULL) + link
 *   backward           dS[b] = upstream * coef[b] * gsym[b] . S[b]  -- gp_scale_rows_batch makes the scaled copy
 *                      of S (fp32 and/or bf16 operand, zero-padded to cols_pad), then one GEMM.
 * Row entropy  L_E = (1/sum_b n_b) sum_{b, n<n_b} -sum_k S log(S + 1e-7):
 *   gp_entropy_fwd     partial[gp_entropy_partials(B, N)] <- per-block sums (finish with gp_loss_finalize)
 *   gp_entropy_bwd     ds (+)= upstream * scale * -(log(S+eps) + S/(S+eps)) on real rows
 * gp_add_scaled: total = (base ? *base : 0) + w * (*term)   (device scalars)
 * ------------------------------------------------------------------------------------------- */
int gp_frob_link_fwd(const float* s, const float* adj, const int32_t* nb, int B, int N, int K,
                     float* partial, float* gsym, gp_stream_t stream);
int gp_frob_finalize(const float* partial, int per_graph, int B, const float* ce, float* total, float* link,
                     float* norm, float* coef, gp_stream_t stream);
int gp_scale_rows_batch(const float* x, const float* scale, const float* upstream, int B, int rows_per_batch,
                        int cols, float* out, long long ldo, void* out_bf16, long long ldob, int cols_pad,
                        gp_stream_t stream);
int gp_entropy_partials(int B, int N);
int gp_entropy_fwd(const float* s, const int32_t* nb, int B, int N, int K, float* partial, gp_stream_t stream);
int gp_entropy_bwd(const float* s, const int32_t* nb, int B, int N, int K, const float* upstream, float scale,
                   float* ds, int accumulate, gp_stream_t stream);
int gp_add_scaled(const float* base, const float* term, float w, float* total, gp_stream_t stream);
/* Device-resident normalisers (node counts known on the device only, e.g. inside a captured CUDA graph):
 *   gp_nb_stats     out[0] = 1 / sum_b nb[b]^2 (encoders.py:1326) ; out[1] = 1 / sum_b nb[b]
 *   gp_mul_add_dev  prod = (*a) * (*b) ; total = (c ? *c : 0) + prod      (all device scalars; outputs optional) */
int gp_nb_stats(const int32_t* nb, int B, float* out, gp_stream_t stream);
/* Optimiser step of train.py:209-210 over flat buffers (dp.FlatAdam):
 *   gp_sumsq_f32      *out = sum_i g[i]^2 (deterministic two-stage; ws: 1024 floats)
 *   gp_adam_step_f32  g' = grad_scale * g (data parallel: 1 / world after the SUM all-reduce, else 1), then
 *                     clip_grad_norm folded into Adam: g'' = g' * min(1, max_norm / (grad_scale * sqrt(*sumsq) + 1e-6))
 *                     (skipped when max_norm <= 0 or sumsq == NULL; *sumsq is the sum of squares of the UNSCALED g);
 *                     torch.optim.Adam's update with t = *step_dev + 1 (no weight decay, no amsgrad); then
 *                     *step_dev += 1.  step_dev: device float counter (CUDA-graph friendly).
 *   gp_clip_scale_f32 the same scaling applied to g in place (for optimisers other than FlatAdam)
 *   gp_multi_axpy_f32 dst_e[i] += alpha * src_e[i] for `count` entries (host array) in ceil(count/64) launches: all
 *                     parameter gradients of one backward pass are added into the flat gradient buffer at once. */
typedef struct gp_axpy_entry { const float* src; float* dst; long long n; } gp_axpy_entry;
int gp_sumsq_f32(const float* g, long long n, float* out, float* ws, gp_stream_t stream);
int gp_adam_step_f32(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                     float eps, float* step_dev, const float* sumsq_dev, float max_norm, float grad_scale,
                     gp_stream_t stream);
int gp_clip_scale_f32(float* g, long long n, const float* sumsq_dev, float max_norm, float pre_scale,
                      gp_stream_t stream);
int gp_multi_axpy_f32(const gp_axpy_entry* entries, int count, float alpha, gp_stream_t stream);
int gp_mul_add_dev(const float* a, const float* b, const float* c, float* total, float* prod, gp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Cross entropy (encoders.py:1127): loss = mean_b -log softmax(logits)[label]; probs saved.
 * Backward: dlogits = upstream * (probs - onehot) / B.
 * ------------------------------------------------------------------------------------------- */
int gp_ce_fwd(const float* logits, const int64_t* label, int B, int C, float* loss, float* probs,
              gp_stream_t stream);
int gp_ce_bwd(const float* probs, const int64_t* label, const float* upstream, int B, int C,
              float* dlogits, gp_stream_t stream);

/* colsum: out[c] (+)= sum_r x[r, c];  ws >= 256*d floats.  relu_mask_bwd: dx = dy * [y > 0]. */
int gp_colsum_f32(const float* x, long long rows, int d, long long ld, float* out, int accumulate,
                  float* ws, gp_stream_t stream);
int gp_relu_mask_bwd(const float* dy, const float* y, long long n, float* dx, gp_stream_t stream);
/* in place: v (+bias) -> v / max(||v||_2, 1e-12) per row (encoders.py:323-326); rnorm[r] = the divisor */
int gp_bias_normalize_f32(float* v, const float* bias, float* rnorm, long long rows, int d, long long ld,
                          int normalize, gp_stream_t stream);
/* ---------------------------------------------------------------------------------------------
 * Set2Set readout of GcnSet2SetEncoder (set2set.py:33-57, called from encoders.py:1144-1157).
 * E [B,N,ldE] node embeddings (d columns; rows n >= nb[b] count as zero rows, the mask of encoders.py:1080),
 * one-layer LSTM weights in torch.nn.LSTM layout (w_ih [4d,2d], w_hh [4d,d], biases [4d] or NULL, gates i f g o).
 * One CTA per graph runs the N sequential steps.  Forward outputs (all saved for the backward):
 *   qs [B,N+1,2d] q*_t (row 0 zeros, row N = the readout), gates [B,N,4d], cells [B,N,d], att [B,N,N].
 * Backward: dout = gradient of qs[:,N,:] (row stride lddout); emits dz [B,N+1,4d] (gate pre-activation gradients,
 * row N zero), dr [B,N,d], de [B,N,N]; the caller forms dW_ih = dz^T qs, dW_hh = dz^T qs[:, :d],
 * db = colsum(dz) and dE_b = att_b^T dr_b + de_b^T qs_b[1:, :d] with gp_bgemm_f32 / gp_colsum_f32.
 * Needs (N + 16 d + 64) floats of shared memory (<= 200 KB).
 * ------------------------------------------------------------------------------------------- */
int gp_set2set_fwd(const float* E, long long ldE, const int32_t* nb, int B, int N, int d, const float* w_ih,
                   const float* w_hh, const float* b_ih, const float* b_hh, float* qs, float* gates, float* cells,
                   float* att, gp_stream_t stream);
int gp_set2set_bwd(const float* E, long long ldE, const int32_t* nb, int B, int N, int d, const float* w_ih,
                   const float* w_hh, const float* gates, const float* cells, const float* att, const float* dout,
                   long long lddout, float* dz, float* dr, float* de, gp_stream_t stream);

/* x[i] = v ;  y[i] += a*x[i]  (buffer initialisation / gradient accumulation for num_pooling >= 2) */
int gp_fill_f32(float* x, long long n, float v, gp_stream_t stream);
int gp_axpy_f32(const float* x, float* y, long long n, float a, gp_stream_t stream);
int gp_fill_i32(int32_t* x, long long n, int32_t v, gp_stream_t stream);
/* Dropout of a GraphConv input (encoders.py:316-317; conv_block layers, training mode): y = keep ? x/(1-p) : 0,
 * fp32 and/or bf16 output, keep = hash(seed, row*d + col) >= p.  Reproducible from the seed: the backward calls it
 * in place on the input gradient (x == y) with the forward's seed.  The mask stream is this library's own -- it
 * cannot reproduce torch's generator, only nn.Dropout's distribution. */
int gp_dropout_f32(const float* x, long long ldx, long long rows, int d, float p, unsigned long long seed,
                   float* y, long long ldy, void* y_bf16, long long ldyb, gp_stream_t stream);
/* dst [rows_dst, cols_dst] = src [rows, cols] in the top-left corner, `fill` elsewhere (padded parameter / gradient
 * copies when the assignment width is padded to a multiple of 8) */
int gp_pad_copy_f32(const float* src, long long ld_src, long long rows, int cols, float* dst, long long ld_dst,
                    long long rows_dst, int cols_dst, float fill, gp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * PACKED schedule for ENZYMES-sized graphs (N <= 128; SURVEY 8(d) small-graph regime, encoders.py:1054-1081,
 * 1231-1300 with num_pooling == 1).  Only the real n_b rows of every graph exist ("packed rows": graph g owns
 * rows rowptr[g] .. rowptr[g+1]); a pad row of a layer is the one vector normalize(bias) and enters the BatchNorm
 * statistics and the bias gradient analytically through cnt_pad[n] = #graphs with n_b <= n
 * (tests/test_pad_row_math_cpu.py, tests/packed_blueprint.py).  The level-0 adjacency is read ONCE per step and
 * kept as per-row neighbour lists (out: A row i, in: A column i), the pooled level keeps its dense [K,K] A'.
 * A launch = one phase for ALL graphs; CTAs walk "tiles" (groups of whole graphs, ~64 packed rows) so that every
 * graph-structured product (A.H, S^T Z, S^T A S, A^T dU) is local to a CTA's shared memory and the dense products
 * (U.W, dV.W^T, U^T dV) run over all rows of the tile at once.  The only cross-graph coupling, BatchNorm per node
 * index (encoders.py:1048-1052), is a sum per node index that a phase accumulates and the next phase consumes:
 * the launch boundary is the grid-wide dependency.
 * ------------------------------------------------------------------------------------------- */
#define GP_PK_ELL 8        /* neighbour-list entries per row kept at a fixed position (longer lists overflow) */
typedef struct gp_pk_tiling {
  const int32_t* rowptr;   /* [B+1] first packed row of each graph; NULL: every graph has `nfix` rows */
  const int32_t* subs;     /* [nsub][4] {first packed row, rows, first graph, end graph}: runs of whole graphs with at
                              most max_rows rows, built by gp_pk_prepare; NULL: `gpt` graphs per run */
  const int32_t* nsub;     /* device scalar; NULL: ceil(B / gpt) */
  const int32_t* rowmeta;  /* [R][2] {node index, graph} of every packed row; NULL: computed from nfix */
  int B, nfix, gpt, max_rows;   /* max_rows: upper bound of the rows of one run (sizes the shared memory) */
} gp_pk_tiling;

typedef struct gp_pk_adj {
  const int32_t* info;     /* [R][2] {first overflow entry, degree} of every packed row; NULL: dense */
  const int32_t* ell;      /* [R][GP_PK_ELL][2] first entries of every row {node index within the graph, value as
                              float bits}, zero padded */
  const int32_t* entries;  /* overflow entries of rows with more than GP_PK_ELL neighbours */
  const float* dense;      /* [B, nfix, nfix] when info == NULL */
  int transposed;          /* dense only: row i of the operator is column i of `dense` */
} gp_pk_adj;

/* packed rows of one layer's normalised output Y (or of the packed input) and the ReLU + BatchNorm applied on read */
typedef struct gp_pk_src {
  const float* y; long long ld; int d;
  const double* sums;      /* [2N] sum relu(y), sum relu(y)^2 per node index over the REAL rows; NULL: rows as they are */
  const float* bias;       /* bias of the producing layer (its pad rows are normalize(bias)); NULL: zero pad rows */
} gp_pk_src;

/* upstream gradient of a layer's output: dense packed rows and/or the max-readout scatter */
typedef struct gp_pk_grad {
  const float* dense; long long ld; int coff;     /* dense[r*ld + coff + c]; NULL: none */
  const float* dout; const int32_t* arg; long long ldo; int ooff;
                           /* dout[g*ldo + ooff + c] lands on node arg[g*ldo + ooff + c] of graph g; NULL: none */
} gp_pk_grad;

typedef struct gp_pk_stack_fwd {
  gp_pk_src in;
  const float* W; const float* b; int dout;     /* W [din, dout] row-major (GraphConv.weight), b [dout] or NULL */
  float* y; float* rnorm;                       /* [R, dout], [R]: Y = V / max(||V||, 1e-12), the clamped norm */
  double* sums_out;                             /* [2N] accumulated (zero on entry); NULL: last layer of the stack */
} gp_pk_stack_fwd;
typedef struct gp_pk_layer_fwd_args {
  gp_pk_tiling tl; gp_pk_adj adj; const float* cnt_pad; int N; int ns; gp_pk_stack_fwd s[2];
} gp_pk_layer_fwd_args;

typedef struct gp_pk_stack_bwd {
  gp_pk_src in;            /* the layer's input, as in the forward */
  gp_pk_src out;           /* this layer's Y with ITS BatchNorm sums (sums NULL: last layer) */
  const float* rnorm;
  const double* msums;     /* [2N] sum gl, sum gl*H of this layer over the real rows (NULL with the last layer) */
  gp_pk_grad gl;           /* upstream gradient of this layer's output */
  const float* W; const float* b; int dout;
  float* dW; float* db;    /* [din, dout], [dout]: ACCUMULATED with atomics (zero on entry); db may be NULL */
  int need_dx;             /* gl_prev = gz_prev + A^T (dV W^T) */
  gp_pk_grad gz_prev; float* gl_prev;           /* [R, din] */
  double* msums_prev;      /* [2N] accumulated sum gl_prev, sum gl_prev*Hin; NULL: the input has no BatchNorm */
  float* dadj; int dadj_acc;                    /* dense level only: dA[g] (+)= dU Hin^T, [B, nfix, nfix] */
} gp_pk_stack_bwd;
typedef struct gp_pk_layer_bwd_args {
  gp_pk_tiling tl; gp_pk_adj adj; gp_pk_adj adj_in; const float* cnt_pad; int N; int ns; gp_pk_stack_bwd s[2];
} gp_pk_layer_bwd_args;

#define GP_PK_MAX_LAYERS 6
typedef struct gp_pk_concat { int L; int F; gp_pk_src slot[GP_PK_MAX_LAYERS]; } gp_pk_concat;

typedef struct gp_pk_pool_args {
  gp_pk_tiling tl; gp_pk_adj adj; gp_pk_adj adj_in; const float* cnt_pad; const int32_t* nb; int N; int K;
  gp_pk_concat z, za;                       /* embedding / assignment concat (encoders.py:1078) */
  const float* Wp; const float* bp;         /* assign_pred: Linear [K, Fa], [K] or NULL (encoders.py:1272) */
  float* S;                                 /* [B, N, K] dense, pad rows zero (encoders.py:1273-1275) */
  float* xp; float* ap;                     /* X' = S^T Z [B,K,F], A' = S^T A S [B,K,K] (encoders.py:1278-1279) */
  float* out; int32_t* arg; long long ldo;  /* level-0 max readout (encoders.py:1257): [B, ldo], first F columns */
  /* backward only */
  const float* dxp; const float* dap;       /* [B,K,F], [B,K,K] */
  const float* dS_ext;                      /* [B,N,K] gradient of S from the losses, or NULL */
  const float* dout;                        /* [B, ldo] gradient of the readout */
  float* gz; float* gza;                    /* [R, F], [R, Fa]: gradients of the two concats */
  float* dWp; float* dbp;                   /* accumulated with atomics (zero on entry) */
} gp_pk_pool_args;

/* nb [B] -> rowptr [B+1], cnt_pad [N], the run table subs (graphs are grouped by windows of `window` packed rows --
 * the graphs whose first row falls into a window -- and every window is split greedily into runs of whole graphs with
 * at most max_rows >= max(N, window) rows; at most 4 runs per window), meta = {R, nsub, 0, 0} (meta[2], meta[3]: the
 * overflow cursors of gp_pk_build_lists).  nb == NULL: every graph has N nodes.  subs: [4 * (B*N / window + 2)][4]. */
int gp_pk_prepare(const int32_t* nb, int B, int N, int window, int max_rows, int32_t* rowptr, float* cnt_pad,
                  int32_t* subs, int32_t* meta, gp_stream_t stream);
/* dense fp32 adjacency [B,N,N] -> neighbour lists of the n_b x n_b blocks (out: rows, in: columns) in ELL + overflow
 * form, rowmeta [R][2], and the packed copies of the inputs: xpack [R, ldxp] <- x [B,N,D] (and axpack <- ax when the
 * assignment GCN has its own features; NULL otherwise).  cursors: two zeroed device ints, `capacity` overflow entries
 * per direction (sum max(deg - GP_PK_ELL, 0) <= sum n_b^2). */
int gp_pk_build_lists(const float* adj, const int32_t* nb, const int32_t* rowptr, int B, int N, int32_t* info_out,
                      int32_t* ell_out, int32_t* ovf_out, int32_t* info_in, int32_t* ell_in, int32_t* ovf_in,
                      int32_t* cursors, long long capacity, int32_t* rowmeta, const float* x, int D, float* xpack,
                      long long ldxp, const float* ax, int Da, float* axpack, long long ldaxp, gp_stream_t stream);
int gp_pk_layer_fwd(const gp_pk_layer_fwd_args* a, gp_stream_t stream);
int gp_pk_layer_bwd(const gp_pk_layer_bwd_args* a, gp_stream_t stream);
int gp_pk_pool_fwd(const gp_pk_pool_args* a, gp_stream_t stream);
int gp_pk_pool_bwd(const gp_pk_pool_args* a, gp_stream_t stream);
/* max readout of a packed concat over each graph's rows (pooled level: encoders.py:1287) */
int gp_pk_readout(const gp_pk_tiling* tl, const gp_pk_concat* z, const float* cnt_pad, const int32_t* nb, int N,
                  float* out, int32_t* arg, long long ldo, int ooff, gp_stream_t stream);
/* link-prediction loss over the real n_b x n_b blocks without materialising P or dl/dP (encoders.py:1311-1331):
 * fwd adds sum -[a log(P+eps) + (1-a) log(1-P+eps)] into the zeroed double *sum; bwd writes
 * dS = alpha * alpha_dev[0] * alpha_dev2[0] * (G + G^T) S (device scalars optional), pad rows zero */
int gp_pk_link_fwd(const float* S, const float* adj, const int32_t* nb, int B, int N, int K, double* sum,
                   gp_stream_t stream);
int gp_pk_link_bwd(const float* S, const float* adj, const int32_t* nb, int B, int N, int K, float alpha,
                   const float* alpha_dev, const float* alpha_dev2, float* dS, gp_stream_t stream);
/* total [1] = base[0] (or 0) + scale * scale_dev[0] (or 1) * sum[0]; link [1] = the second term */
int gp_pk_link_finalize(const double* sum, double scale, const float* scale_dev, const float* base, float* total,
                        float* link, gp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GP_B200_H */
