// train.cu — train-mode kernels of the point-wise layers for sm_100a.
//
// The reference trains through EquivariantLayer = Conv1d(k=1) -> MyBatchNorm1d (batch statistics,
// models/layers.py:22-70) -> ReLU, and back-propagates through the per-node pool as the backward
// of a gather (models/networks.py:185). PyTorch runs that as conv + batch_norm + relu (+ three
// backward kernels each) and a zero-fill + scatter for the gather. Here:
//
//   bn_stats / bn_apply          batch statistics of x[B,C,P] per channel, then
//                                y = relu(gamma * (x - mean) * invstd + beta) in ONE elementwise pass
//   bn_backward (2 launches)     per-channel sums of g and g*xhat (g = dy gated by the ReLU of the
//                                recomputed output), then dx = gamma*invstd*(g - mean(g) -
//                                xhat*mean(g*xhat)); dgamma, dbeta fall out of the sums
//   index_max_backward           grad of first_pn_out through the masked gather: zero-fill +
//                                scatter-add of grad[b,c,k] at idx[b,c,k] (duplicates — every empty
//                                node gathers point 0 — are added by ONE thread in ascending k)
//   tc_pack_device               fp32 weights -> tcgen05 hi/lo blob ON THE DEVICE (weights change
//                                every optimizer step; the eval path packs on the host once), also
//                                transposed: the dgrad GEMM dx = W^T dy runs on the same generic
//                                tcgen05 layer kernel as the forward (csrc/pointwise_tc.cu)
//
// All reductions are two-stage with fp64 partial sums in a fixed order: bit-reproducible.
#include <algorithm>
#include <cmath>

#include <cuda_fp16.h>

#include "common.cuh"

namespace sonet {

constexpr int BN_THREADS = 256;

// stage 1: CTA (c, s) sums the clouds b = s, s+S, ... of channel c. want_gx: sums of g and g*xhat
// for the backward (g = dy * [gamma*xhat + beta > 0] when relu), else sums of x and x*x.
template <bool BWD>
__global__ void __launch_bounds__(BN_THREADS)
    bn_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                      const float* __restrict__ mean, const float* __restrict__ invstd,
                      const float* __restrict__ gamma, const float* __restrict__ beta, int B, int C,
                      int P, int relu, int S, double* __restrict__ partial) {
  const int c = blockIdx.x, s = blockIdx.y;
  double a0 = 0.0, a1 = 0.0;
  float mu = 0.f, is = 1.f, ga = 1.f, be = 0.f;
  if (BWD) {
    mu = mean[c];
    is = invstd[c];
    ga = gamma ? gamma[c] : 1.f;
    be = beta ? beta[c] : 0.f;
  }
  for (int b = s; b < B; b += S) {
    const float* xr = x + (static_cast<size_t>(b) * C + c) * P;
    const float* gr = BWD ? dy + (static_cast<size_t>(b) * C + c) * P : nullptr;
    float f0 = 0.f, f1 = 0.f;            // fp32 within one row chunk, fp64 across
    int cnt = 0;
    for (int p = threadIdx.x; p < P; p += BN_THREADS) {
      const float v = xr[p];
      if (BWD) {
        const float xh = (v - mu) * is;
        float g = gr[p];
        if (relu && !(fmaf(ga, xh, be) > 0.f)) g = 0.f;
        f0 += g;
        f1 = fmaf(g, xh, f1);
      } else {
        f0 += v;
        f1 = fmaf(v, v, f1);
      }
      if (++cnt == 64) {                 // flush to fp64 every 64 terms
        a0 += f0; a1 += f1; f0 = f1 = 0.f; cnt = 0;
      }
    }
    a0 += f0;
    a1 += f1;
  }
  __shared__ double r0[BN_THREADS], r1[BN_THREADS];
  r0[threadIdx.x] = a0;
  r1[threadIdx.x] = a1;
  __syncthreads();
  for (int w = BN_THREADS / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) {
      r0[threadIdx.x] += r0[threadIdx.x + w];
      r1[threadIdx.x] += r1[threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[(static_cast<size_t>(c) * S + s) * 2] = r0[0];
    partial[(static_cast<size_t>(c) * S + s) * 2 + 1] = r1[0];
  }
}

__global__ void bn_stats_final_kernel(const double* __restrict__ partial, int C, int S, double n,
                                      float eps, float* __restrict__ mean,
                                      float* __restrict__ var, float* __restrict__ invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0;
  for (int s = 0; s < S; ++s) {
    s0 += partial[(static_cast<size_t>(c) * S + s) * 2];
    s1 += partial[(static_cast<size_t>(c) * S + s) * 2 + 1];
  }
  const double m = s0 / n;
  double v = s1 / n - m * m;
  if (v < 0.0) v = 0.0;
  mean[c] = static_cast<float>(m);
  var[c] = static_cast<float>(v);                                  // biased (F.batch_norm normalises with it)
  invstd[c] = static_cast<float>(1.0 / sqrt(v + static_cast<double>(eps)));
}

__global__ void bn_bwd_final_kernel(const double* __restrict__ partial, int C, int S,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0;
  for (int s = 0; s < S; ++s) {
    s0 += partial[(static_cast<size_t>(c) * S + s) * 2];
    s1 += partial[(static_cast<size_t>(c) * S + s) * 2 + 1];
  }
  dbeta[c] = static_cast<float>(s0);
  dgamma[c] = static_cast<float>(s1);
}

// y = act(gamma*(x-mean)*invstd + beta); grid (ceil(P/1024), C, B)
__global__ void __launch_bounds__(256)
    bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, int C, int P, int relu, float* __restrict__ y) {
  const int c = blockIdx.y, b = blockIdx.z;
  const float is = invstd[c], ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f, mu = mean[c];
  const size_t o = (static_cast<size_t>(b) * C + c) * P;
  for (int p = blockIdx.x * 1024 + threadIdx.x; p < min(P, (static_cast<int>(blockIdx.x) + 1) * 1024);
       p += 256) {
    const float v = fmaf(ga, (x[o + p] - mu) * is, be);
    y[o + p] = relu ? fmaxf(v, 0.f) : v;
  }
}

// dx = gamma*invstd*(g - dbeta/n - xhat*dgamma/n)
__global__ void __launch_bounds__(256)
    bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                        const float* __restrict__ mean, const float* __restrict__ invstd,
                        const float* __restrict__ gamma, const float* __restrict__ beta,
                        const float* __restrict__ dgamma, const float* __restrict__ dbeta, int C,
                        int P, int relu, float inv_n, float* __restrict__ dx) {
  const int c = blockIdx.y, b = blockIdx.z;
  const float is = invstd[c], ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f, mu = mean[c];
  const float k0 = dbeta[c] * inv_n, k1 = dgamma[c] * inv_n;
  const size_t o = (static_cast<size_t>(b) * C + c) * P;
  for (int p = blockIdx.x * 1024 + threadIdx.x; p < min(P, (static_cast<int>(blockIdx.x) + 1) * 1024);
       p += 256) {
    const float xh = (x[o + p] - mu) * is;
    float g = dy[o + p];
    if (relu && !(fmaf(ga, xh, be) > 0.f)) g = 0.f;
    dx[o + p] = ga * is * (g - k0 - xh * k1);
  }
}

// grad_data[b,c,idx[b,c,k]] += grad_out[b,c,k], k ascending, one thread per (b,c) row
__global__ void __launch_bounds__(128)
    index_max_backward_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ idx,
                              long long rows, int N, int K, float* __restrict__ grad_data) {
  const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float* g = grad_data + r * N;
  const float* go = grad_out + r * K;
  const int32_t* ix = idx + r * K;
  for (int k = 0; k < K; ++k) {
    const int n = min(max(ix[k], 0), N - 1);
    g[n] += go[k];
  }
}

// ---- device-side weight packing for the tcgen05 layer ---------------------------------------------
__global__ void __launch_bounds__(256)
    absmax_kernel(const float* __restrict__ w, long long n, unsigned* __restrict__ out_bits) {
  float m = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    m = fmaxf(m, fabsf(w[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));   // non-negative floats
}
// scale2[0] = power-of-two pre-scale with max|scale*W| in [256,512), scale2[1] = 1/scale
__global__ void scale_from_absmax_kernel(const unsigned* __restrict__ bits, float* __restrict__ scale2) {
  const float mx = __uint_as_float(*bits);
  float sc = 1.f;
  if (mx > 0.f && isfinite(mx)) {
    int e;
    frexpf(mx, &e);
    sc = ldexpf(1.f, 9 - e);
  }
  scale2[0] = sc;
  scale2[1] = 1.f / sc;
}
// blob layout of csrc/pointwise_tc.cu (tc_pack_matrix): n-tiles of <= 256 rows, 64-channel K chunks,
// two [nw x 32] stages per chunk, hi image then lo image, 8x16-byte core matrices.
__global__ void __launch_bounds__(256)
    tc_pack_device_kernel(const float* __restrict__ W, int Cout, int Cin, int transpose,
                          const float* __restrict__ scale2, unsigned char* __restrict__ blob,
                          int cpad, int kch) {
  // logical matrix Wl[n][k], n < Nl, k < Kl: Wl = W (Nl=Cout,Kl=Cin) or W^T (Nl=Cin,Kl=Cout)
  const int Nl = transpose ? Cin : Cout, Kl = transpose ? Cout : Cin;
  const long long total = static_cast<long long>(cpad) * kch * 64;
  const float sc = scale2[0];
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(t % (kch * 64));
    const int n = static_cast<int>(t / (kch * 64));
    const int nt = n / 256, r = n - nt * 256;
    const int nw = min(256, cpad - nt * 256);
    const int kc = k >> 6, hh = (k >> 5) & 1, kk = k & 31;
    float w = 0.f;
    if (n < Nl && k < Kl) w = (transpose ? W[static_cast<size_t>(k) * Cin + n]
                                         : W[static_cast<size_t>(n) * Cin + k]) * sc;
    const __half h = __float2half_rn(w);
    const __half l = __float2half_rn(w - __half2float(h));
    // tile nt starts after nt full-width tiles: nt * 256 rows * kch chunks * 64 k * 4 B
    const size_t stage = static_cast<size_t>(nt) * 256 * kch * 64 * 4 +
                         (static_cast<size_t>(kc) * 2 + hh) * nw * 32 * 4;
    const uint32_t o = (r >> 3) * 512 + (kk >> 3) * 128 + (r & 7) * 16 + (kk & 7) * 2;
    *reinterpret_cast<__half*>(blob + stage + o) = h;
    *reinterpret_cast<__half*>(blob + stage + static_cast<size_t>(nw) * 32 * 2 + o) = l;
  }
}

// ---- wgrad helpers ------------------------------------------------------------------------------------
// dyT[k = b*P + p][co] = dy[b][co][p] * s  (s = scale2[0], a power of two): the "activation" operand
// of the wgrad GEMM in the [K][rows] layout the generic tcgen05 layer consumes. 32x32 smem tiles.
__global__ void __launch_bounds__(256)
    transpose_scale_kernel(const float* __restrict__ dy, int C, int P, const float* __restrict__ scale2,
                           float* __restrict__ dyT) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float s = scale2[0];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    tile[i][tx] = (c < C && p < P) ? dy[(static_cast<size_t>(b) * C + c) * P + p] * s : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    if (p < P && c < C) dyT[(static_cast<size_t>(b) * P + p) * C + c] = tile[tx][i];
  }
}
// out2[0] = a[1] * b[1] (product of two inverse scales); out2[1] unused
__global__ void inv_product_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                   float* __restrict__ out) {
  out[0] = a[1] * b[1];
}
// x [B][Cin][P] as the "weight" matrix Wl[n = ci][k = b*P + p] of the wgrad GEMM, packed into the
// tcgen05 blob layout (see tc_pack_device_kernel); K padded with zeros up to kch*64.
__global__ void __launch_bounds__(256)
    tc_pack_points_kernel(const float* __restrict__ x, int B, int Cin, int P,
                          const float* __restrict__ scale2, unsigned char* __restrict__ blob, int cpad,
                          int kch) {
  const long long K = static_cast<long long>(B) * P;
  const long long total = static_cast<long long>(cpad) * kch * 64;
  const float sc = scale2[0];
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    // consecutive threads -> consecutive k of one row n: coalesced reads of x along p
    const long long k = t % (static_cast<long long>(kch) * 64);
    const int n = static_cast<int>(t / (static_cast<long long>(kch) * 64));
    const int nt = n / 256, r = n - nt * 256;
    const int nw = min(256, cpad - nt * 256);
    const long long kc = k >> 6;
    const int hh = static_cast<int>((k >> 5) & 1), kk = static_cast<int>(k & 31);
    float w = 0.f;
    if (n < Cin && k < K) {
      const long long b = k / P;
      const int p = static_cast<int>(k - b * P);
      w = x[(static_cast<size_t>(b) * Cin + n) * P + p] * sc;
    }
    const __half h = __float2half_rn(w);
    const __half l = __float2half_rn(w - __half2float(h));
    const size_t stage = static_cast<size_t>(nt) * 256 * kch * 64 * 4 +
                         (static_cast<size_t>(kc) * 2 + hh) * nw * 32 * 4;
    const uint32_t o = (r >> 3) * 512 + (kk >> 3) * 128 + (r & 7) * 16 + (kk & 7) * 2;
    *reinterpret_cast<__half*>(blob + stage + o) = h;
    *reinterpret_cast<__half*>(blob + stage + static_cast<size_t>(nw) * 32 * 2 + o) = l;
  }
}

}  // namespace sonet

static int bn_splits(int B, int C) {
  int S = std::max(1, std::min(B, (4 * sonet::sm_count() + C - 1) / C));
  return std::min(S, 64);
}

extern "C" int sonet_bn_partial_slots(int B, int C) { return bn_splits(B, C) * C * 2; }

extern "C" int sonet_bn_train_forward_f32(const float* x, const float* gamma, const float* beta,
                                          int B, int C, int P, float eps, int relu, double* partial,
                                          float* y, float* save_mean, float* save_var,
                                          float* save_invstd, sonet_stream_t stream) {
  using namespace sonet;
  SONET_REQUIRE(B >= 1 && C >= 1 && P >= 1 && B <= 65535 && C <= 65535, "bn_train_forward: bad dims");
  SONET_REQUIRE(x && partial && y && save_mean && save_var && save_invstd, "bn_train_forward: null pointer");
  cudaStream_t st = as_stream(stream);
  const int S = bn_splits(B, C);
  bn_partial_kernel<false><<<dim3(C, S), BN_THREADS, 0, st>>>(x, nullptr, nullptr, nullptr, nullptr,
                                                             nullptr, B, C, P, 0, S, partial);
  bn_stats_final_kernel<<<(C + 127) / 128, 128, 0, st>>>(partial, C, S, static_cast<double>(B) * P, eps,
                                                        save_mean, save_var, save_invstd);
  bn_apply_kernel<<<dim3((P + 1023) / 1024, C, B), 256, 0, st>>>(x, save_mean, save_invstd, gamma, beta,
                                                               C, P, relu, y);
  return check_launch("bn_train_forward");
}

extern "C" int sonet_bn_train_backward_f32(const float* dy, const float* x, const float* mean,
                                           const float* invstd, const float* gamma,
                                           const float* beta, int B, int C, int P, int relu,
                                           double* partial, float* dx, float* dgamma, float* dbeta,
                                           sonet_stream_t stream) {
  using namespace sonet;
  SONET_REQUIRE(B >= 1 && C >= 1 && P >= 1 && B <= 65535 && C <= 65535, "bn_train_backward: bad dims");
  SONET_REQUIRE(dy && x && mean && invstd && partial && dx && dgamma && dbeta,
                "bn_train_backward: null pointer");
  cudaStream_t st = as_stream(stream);
  const int S = bn_splits(B, C);
  bn_partial_kernel<true><<<dim3(C, S), BN_THREADS, 0, st>>>(x, dy, mean, invstd, gamma, beta, B, C, P,
                                                            relu, S, partial);
  bn_bwd_final_kernel<<<(C + 127) / 128, 128, 0, st>>>(partial, C, S, dgamma, dbeta);
  bn_bwd_apply_kernel<<<dim3((P + 1023) / 1024, C, B), 256, 0, st>>>(
      dy, x, mean, invstd, gamma, beta, dgamma, dbeta, C, P, relu,
      static_cast<float>(1.0 / (static_cast<double>(B) * P)), dx);
  return check_launch("bn_train_backward");
}

extern "C" int sonet_index_max_backward_f32(const float* grad_out, const int32_t* idx, int B, int C,
                                            int N, int K, float* grad_data, sonet_stream_t stream) {
  using namespace sonet;
  SONET_REQUIRE(B >= 0 && C >= 0 && N >= 1 && K >= 1, "index_max_backward: bad dims");
  const long long rows = static_cast<long long>(B) * C;
  if (rows == 0) return SONET_OK;
  SONET_REQUIRE(grad_out && idx && grad_data, "index_max_backward: null pointer");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(grad_data, 0, sizeof(float) * static_cast<size_t>(rows) * N, st);
  index_max_backward_kernel<<<static_cast<unsigned>((rows + 127) / 128), 128, 0, st>>>(
      grad_out, idx, rows, N, K, grad_data);
  return check_launch("index_max_backward");
}

extern "C" long long sonet_pointwise_tc_blob_bytes(int Cout, int Cin);

extern "C" int sonet_pointwise_tc_pack_device(const float* W, int Cout, int Cin, int transpose,
                                              void* blob, float* scale2, unsigned* scratch_bits,
                                              sonet_stream_t stream) {
  using namespace sonet;
  SONET_REQUIRE(W && blob && scale2 && scratch_bits && Cout >= 1 && Cin >= 1,
                "pointwise_tc_pack_device: bad args");
  cudaStream_t st = as_stream(stream);
  const long long n = static_cast<long long>(Cout) * Cin;
  cudaMemsetAsync(scratch_bits, 0, sizeof(unsigned), st);
  absmax_kernel<<<static_cast<unsigned>(std::min<long long>((n + 255) / 256, 1024)), 256, 0, st>>>(
      W, n, scratch_bits);
  scale_from_absmax_kernel<<<1, 1, 0, st>>>(scratch_bits, scale2);
  const int Nl = transpose ? Cin : Cout, Kl = transpose ? Cout : Cin;
  const int cpad = (Nl + 63) / 64 * 64;
  const int kch = ((Kl + 15) / 16 * 16 + 63) / 64;
  const long long total = static_cast<long long>(cpad) * kch * 64;
  tc_pack_device_kernel<<<static_cast<unsigned>(std::min<long long>((total + 255) / 256, 4096)), 256, 0,
                          st>>>(W, Cout, Cin, transpose, scale2, static_cast<unsigned char*>(blob),
                                cpad, kch);
  return check_launch("pointwise_tc_pack_device");
}

// power-of-two scale of a device tensor: scale2[0] = s with max|s*t| in [256,512), scale2[1] = 1/s
extern "C" int sonet_absmax_scale_f32(const float* t, long long n, float* scale2, unsigned* scratch_bits,
                                      sonet_stream_t stream) {
  using namespace sonet;
  SONET_REQUIRE(t && scale2 && scratch_bits && n >= 1, "absmax_scale: bad args");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(scratch_bits, 0, sizeof(unsigned), st);
  absmax_kernel<<<static_cast<unsigned>(std::min<long long>((n + 255) / 256, 2048)), 256, 0, st>>>(
      t, n, scratch_bits);
  scale_from_absmax_kernel<<<1, 1, 0, st>>>(scratch_bits, scale2);
  return check_launch("absmax_scale");
}

extern "C" int sonet_pointwise_tc_forward_dev(const float* x0, int C0, int B, int P, const void* blob,
                                              const float* inv_scale_dev, const float* act_scale_dev,
                                              const float* shift, int Cout, int relu, int splits,
                                              float* out, float* scratch, sonet_stream_t stream);

// wgrad of a 1x1 convolution on tcgen05: dW[co][ci] = sum_{b,p} dy[b,co,p] * x[b,ci,p].
// Written as the generic layer GEMM out[rows = co][n = ci] = sum_k A[k][co] * Wl[ci][k] with the
// contraction index k = (b, p): A = dy transposed to [K][Cout] (and pre-scaled: gradients are far
// below fp16's normal range), Wl = x packed as a weight blob, K split over CTAs with a
// deterministic reduce. Result: dWT [Cin][Cout] (the caller views it transposed).
// Scratch (device): dyT  Kpad*Cout floats, blob sonet_wgrad_blob_bytes(), part splits*Cin*Cout floats,
// small: 8 floats + 2 uint32.
extern "C" long long sonet_wgrad_kpad(int B, int P, int splits) {
  const long long K = static_cast<long long>(B) * P;
  const long long q = 64LL * splits;
  return (K + q - 1) / q * q;
}
extern "C" long long sonet_wgrad_blob_bytes(int Cin, long long Kpad) {
  const long long cpad = (Cin + 63) / 64 * 64;
  return cpad * Kpad * 4;
}
extern "C" int sonet_wgrad_tc_f32(const float* dy, const float* x, int B, int Cout, int Cin, int P,
                                  int splits, float* dyT, void* blob, float* part, float* small,
                                  float* dWT, sonet_stream_t stream) {
  using namespace sonet;
  SONET_REQUIRE(dy && x && dyT && blob && part && small && dWT, "wgrad_tc: null pointer");
  SONET_REQUIRE(B >= 1 && P >= 1 && Cout >= 64 && Cout % 64 == 0 && Cin >= 1 && splits >= 1 &&
                    B <= 65535,
                "wgrad_tc: needs Cout %% 64 == 0 (got Cout=%d, Cin=%d)", Cout, Cin);
  cudaStream_t st = as_stream(stream);
  const long long K = static_cast<long long>(B) * P;
  const long long Kpad = sonet_wgrad_kpad(B, P, splits);
  SONET_REQUIRE(Kpad < (1LL << 31), "wgrad_tc: too many points");
  float* s_dy = small;            // [2]
  float* s_x = small + 2;         // [2]
  float* inv = small + 4;         // [1]
  unsigned* bits = reinterpret_cast<unsigned*>(small + 6);
  int rc = sonet_absmax_scale_f32(dy, static_cast<long long>(B) * Cout * P, s_dy, bits, stream);
  if (rc) return rc;
  rc = sonet_absmax_scale_f32(x, static_cast<long long>(B) * Cin * P, s_x, bits + 1, stream);
  if (rc) return rc;
  inv_product_kernel<<<1, 1, 0, st>>>(s_dy, s_x, inv);
  if (Kpad > K)   // zero rows of the padded tail (the matching weight columns are packed as zeros too)
    cudaMemsetAsync(dyT + K * Cout, 0, sizeof(float) * static_cast<size_t>(Kpad - K) * Cout, st);
  transpose_scale_kernel<<<dim3((P + 31) / 32, (Cout + 31) / 32, B), 256, 0, st>>>(dy, Cout, P, s_dy, dyT);
  const int cpad = (Cin + 63) / 64 * 64;
  const int kch = static_cast<int>(Kpad / 64);
  const long long total = static_cast<long long>(cpad) * Kpad;
  tc_pack_points_kernel<<<static_cast<unsigned>(std::min<long long>((total + 255) / 256, 16384)), 256, 0,
                          st>>>(x, B, Cin, P, s_x, static_cast<unsigned char*>(blob), cpad, kch);
  rc = check_launch("wgrad_tc(prepare)");
  if (rc) return rc;
  // activations: dyT viewed as [B'=1][C'=Kpad][P'=Cout] (already scaled: no act pre-scale needed)
  return sonet_pointwise_tc_forward_dev(dyT, static_cast<int>(Kpad), 1, Cout, blob, inv, nullptr, nullptr,
                                        Cin, 0, splits, dWT, part, stream);
}
