// som_sort.cu — node-sorted point order for the fused max-pool path, and its finalisation.
//
// The reference pools the first PointResNet's output per SOM node with index_max on the stacked
// copies in their original order (models/networks.py:181-185). When `first_pn_out` itself is not
// needed (classifier, auto-encoder) sonet_b200 never writes it: the tcgen05 kernel consumes the
// stacked copies GROUPED BY NODE, so that the 32 lanes of an epilogue warp (32 consecutive
// sorted copies) belong to one or two nodes and the per-node max of a channel is one warp `redux`
// plus one atomic max (csrc/pointmlp_tc.cu, POOL variant). This file provides
//
//   som_sort_decenter_kernel   per cloud: bucket the k*N stacked copies by node (offsets from the
//                              node counts, slots handed out by shared-memory atomics — the order
//                              inside a node is arbitrary, max does not care) and write, in sorted
//                              order, the decentred coordinates + normals (models/networks.py:168-172)
//                              and the node id of every sorted position; also the sorted position
//                              of stacked copy 0 (the "point 0" that empty nodes gather).
//   pool_finalize_kernel       ordered-int keys -> values with the reference's semantics
//                              (max must be > -1000, else the feature of point 0,
//                              index_max.cpp:80-81,103 + networks.py:185) and key reset.
#include <algorithm>

#include "common.cuh"

namespace sonet {

constexpr int SORT_THREADS = 1024;
constexpr int SORT_MAX_M = 256;

__global__ void __launch_bounds__(SORT_THREADS)
    som_sort_decenter_kernel(const float* __restrict__ x, const float* __restrict__ sn,
                             const float* __restrict__ cluster_mean,
                             const int32_t* __restrict__ idx32, const int32_t* __restrict__ count,
                             int N, int M, int k, float* __restrict__ x_sorted,
                             int32_t* __restrict__ node_sorted, int32_t* __restrict__ pos0) {
  __shared__ int offs[SORT_MAX_M];
  __shared__ int cursor[SORT_MAX_M];
  __shared__ float cm[3 * SORT_MAX_M];
  const int b = blockIdx.x;
  const int kN = k * N;
  if (threadIdx.x == 0) {
    int run = 0;
    for (int m = 0; m < M; ++m) {
      offs[m] = run;
      run += count[static_cast<size_t>(b) * M + m];
    }
  }
  for (int m = threadIdx.x; m < M; m += SORT_THREADS) cursor[m] = 0;
  for (int i = threadIdx.x; i < 3 * M; i += SORT_THREADS)
    cm[i] = cluster_mean[static_cast<size_t>(b) * 3 * M + i];
  __syncthreads();
  const int CA = sn ? 6 : 3;
  const float* xb = x + static_cast<size_t>(b) * 3 * N;
  const float* sb = sn ? sn + static_cast<size_t>(b) * 3 * N : nullptr;
  float* ob = x_sorted + static_cast<size_t>(b) * CA * kN;
  for (int j = threadIdx.x; j < kN; j += SORT_THREADS) {
    const int node = min(max(idx32[static_cast<size_t>(b) * kN + j], 0), M - 1);
    const int n = j % N;
    const int pos = offs[node] + atomicAdd(&cursor[node], 1);
#pragma unroll
    for (int c = 0; c < 3; ++c)
      ob[static_cast<size_t>(c) * kN + pos] = __fsub_rn(xb[c * N + n], cm[c * M + node]);
    if (sb) {
#pragma unroll
      for (int c = 0; c < 3; ++c) ob[static_cast<size_t>(3 + c) * kN + pos] = sb[c * N + n];
    }
    node_sorted[static_cast<size_t>(b) * kN + pos] = node;
    if (j == 0) pos0[b] = pos;
  }
}

constexpr int POOL_KEY_INIT = static_cast<int>(0x80000000u);  // below the key of every float

__global__ void __launch_bounds__(256)
    pool_finalize_kernel(int32_t* __restrict__ keys, const float* __restrict__ p0, int C, int M,
                         long long total, float* __restrict__ out_val) {
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int key = keys[t];
    keys[t] = POOL_KEY_INIT;  // ready for the next forward
    const int bits = key ^ ((key >> 31) & 0x7fffffff);
    const float v = __int_as_float(bits);
    const long long bc = t / M;
    out_val[t] = (key != POOL_KEY_INIT && v > -1000.0f) ? v : __ldg(p0 + bc);
  }
}

__global__ void __launch_bounds__(256)
    pool_keys_init_kernel(int32_t* __restrict__ keys, long long total) {
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x)
    keys[t] = POOL_KEY_INIT;
}

}  // namespace sonet

extern "C" int sonet_som_sort_decenter(const float* x, const float* sn, const float* cluster_mean,
                                       const int32_t* min_idx_i32, const int32_t* count, int B, int N,
                                       int M, int k, float* x_sorted, int32_t* node_sorted,
                                       int32_t* pos0, sonet_stream_t stream) {
  using namespace sonet;
  SONET_REQUIRE(B >= 0 && N >= 0 && k >= 1, "som_sort_decenter: bad dimension");
  SONET_REQUIRE(M >= 1 && M <= SORT_MAX_M, "som_sort_decenter: M=%d out of range", M);
  if (B == 0 || N == 0) return SONET_OK;
  SONET_REQUIRE(x && cluster_mean && min_idx_i32 && count && x_sorted && node_sorted && pos0,
                "som_sort_decenter: null pointer");
  som_sort_decenter_kernel<<<B, SORT_THREADS, 0, as_stream(stream)>>>(
      x, sn, cluster_mean, min_idx_i32, count, N, M, k, x_sorted, node_sorted, pos0);
  return check_launch("som_sort_decenter");
}

extern "C" int sonet_pool_keys_init(int32_t* keys, long long n, sonet_stream_t stream) {
  using namespace sonet;
  SONET_REQUIRE(n >= 0, "pool_keys_init: negative size");
  if (n == 0) return SONET_OK;
  SONET_REQUIRE(keys, "pool_keys_init: null pointer");
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, 8LL * sm_count()));
  pool_keys_init_kernel<<<grid, 256, 0, as_stream(stream)>>>(keys, n);
  return check_launch("pool_keys_init");
}

extern "C" int sonet_pool_finalize(int32_t* keys, const float* p0, int B, int C, int M,
                                   float* out_val, sonet_stream_t stream) {
  using namespace sonet;
  SONET_REQUIRE(B >= 0 && C >= 0 && M >= 1, "pool_finalize: bad dimension");
  const long long total = static_cast<long long>(B) * C * M;
  if (total == 0) return SONET_OK;
  SONET_REQUIRE(keys && p0 && out_val, "pool_finalize: null pointer");
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 8LL * sm_count()));
  pool_finalize_kernel<<<grid, 256, 0, as_stream(stream)>>>(keys, p0, C, M, total, out_val);
  return check_launch("pool_finalize");
}
