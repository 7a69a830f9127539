// Host-side half of the adjacency feed (no CUDA here): bit-packing of a dense fp32 {0,1} adjacency.
//
// End to end, a training step of the large configuration is bound by PCIe, not by the GPU: the reference's feed hands
// the model a dense fp32 adjacency (train.py:197: 4 bytes per entry, 4.3 GB per 256 x 2048^2 batch = 66 ms of PCIe at
// 65 GB/s against 18 ms of GPU work).  Every entry is 0 or 1, so the host cores can shrink it 32x before the copy:
// 16 threads pack ~87 GB/s on the B200 box, i.e. FASTER than PCIe moves the raw floats, and both can run at once
// (the feed packs a fraction of the graphs while the rest crosses PCIe as fp32; bench.py picks the fraction from the
// two measured rates).  gp_adj_prepare (adjprep.cu) expands the bits to the bf16 operand on the device.
// Exact: the packer reports any entry outside {0,1}, in which case the caller sends those graphs as fp32.
#include <immintrin.h>
#include <stdint.h>
#include <string.h>
#include <thread>
#include <vector>

static int pack_rows_scalar(const float* a, long long rows, int N, uint8_t* out, long long ldb) {
  int bad = 0;
  for (long long r = 0; r < rows; ++r) {
    const float* p = a + r * (long long)N;
    uint8_t* o = out + r * ldb;
    memset(o, 0, (size_t)ldb);
    for (int j = 0; j < N; ++j) {
      const float x = p[j];
      if (x != 0.f) {
        o[j >> 3] |= (uint8_t)(1u << (j & 7));
        if (x != 1.f) bad = 1;
      }
    }
  }
  return bad;
}

__attribute__((target("avx2"))) static int pack_rows_avx2(const float* a, long long rows, int N, uint8_t* out,
                                                          long long ldb) {
  int bad = 0;
  const __m256 zero = _mm256_setzero_ps(), one = _mm256_set1_ps(1.f);
  const int nfull = N >> 3;
  for (long long r = 0; r < rows; ++r) {
    const float* p = a + r * (long long)N;
    uint8_t* o = out + r * ldb;
    for (int q = 0; q < nfull; ++q) {
      const __m256 v = _mm256_loadu_ps(p + 8 * q);
      const int nz = _mm256_movemask_ps(_mm256_cmp_ps(v, zero, _CMP_NEQ_UQ));
      const int is1 = _mm256_movemask_ps(_mm256_cmp_ps(v, one, _CMP_EQ_OQ));
      bad |= nz & ~is1;
      o[q] = (uint8_t)nz;
    }
    for (long long q = nfull; q < ldb; ++q) o[q] = 0;
    for (int j = nfull << 3; j < N; ++j) {
      const float x = p[j];
      if (x != 0.f) {
        o[j >> 3] |= (uint8_t)(1u << (j & 7));
        if (x != 1.f) bad = 1;
      }
    }
  }
  return bad;
}

__attribute__((target("avx512f"))) static int pack_rows_avx512(const float* a, long long rows, int N, uint8_t* out,
                                                               long long ldb) {
  int bad = 0;
  const __m512 zero = _mm512_setzero_ps(), one = _mm512_set1_ps(1.f);
  const int n16 = N >> 4;
  for (long long r = 0; r < rows; ++r) {
    const float* p = a + r * (long long)N;
    uint8_t* o = out + r * ldb;
    for (int q = 0; q < n16; ++q) {
      const __m512 v = _mm512_loadu_ps(p + 16 * q);
      const __mmask16 nz = _mm512_cmp_ps_mask(v, zero, _CMP_NEQ_UQ);
      const __mmask16 is1 = _mm512_cmp_ps_mask(v, one, _CMP_EQ_OQ);
      bad |= (int)(nz & ~is1);
      o[2 * q] = (uint8_t)(nz & 0xff);
      o[2 * q + 1] = (uint8_t)(nz >> 8);
    }
    for (long long q = 2 * n16; q < ldb; ++q) o[q] = 0;
    for (int j = n16 << 4; j < N; ++j) {
      const float x = p[j];
      if (x != 0.f) {
        o[j >> 3] |= (uint8_t)(1u << (j & 7));
        if (x != 1.f) bad = 1;
      }
    }
  }
  return bad;
}

extern "C" int gp_host_pack_adj_bits(const float* adj_host, long long rows, int N, void* out_host, long long ldb,
                                     int threads, int* non01) {
  if (adj_host == nullptr || out_host == nullptr || rows < 0 || N <= 0 || ldb < (N + 7) / 8) return -1;
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  if ((long long)threads > rows) threads = rows > 0 ? (int)rows : 1;
  const bool avx2 = __builtin_cpu_supports("avx2");
  const bool avx512 = __builtin_cpu_supports("avx512f");
  uint8_t* out = static_cast<uint8_t*>(out_host);
  std::vector<int> bad((size_t)threads, 0);
  auto work = [&](int t) {
    const long long lo = rows * t / threads, hi = rows * (t + 1) / threads;
    const float* a = adj_host + lo * (long long)N;
    uint8_t* o = out + lo * ldb;
    bad[(size_t)t] = avx512 ? pack_rows_avx512(a, hi - lo, N, o, ldb)
                            : (avx2 ? pack_rows_avx2(a, hi - lo, N, o, ldb) : pack_rows_scalar(a, hi - lo, N, o, ldb));
  };
  if (threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    pool.reserve((size_t)threads);
    for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
    for (auto& th : pool) th.join();
  }
  int any = 0;
  for (int b : bad) any |= b;
  if (non01 != nullptr) *non01 = any != 0;
  return 0;
}
