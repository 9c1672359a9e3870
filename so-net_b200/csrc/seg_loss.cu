// Segmentation loss of the segmenter wrapper: mean over points of -log_softmax(score)[target].
//
// Replaces CrossEntropyLossSeg.forward (reference models/losses.py:30-43: log_softmax over the
// class axis of [B,classes,N] scores followed by NLLLoss, size_average=True, no class weights) as
// it is evaluated by models/segmenter.py:129-131 (test_model) on every forward. PyTorch runs this
// as a 32-block spatial soft-max plus three small kernels (≈50 µs at B=32, N=1024, 50 classes);
// here: one pass over the scores + a one-block deterministic final sum.
//
//   kernel 1: thread per point (coalesced over n for every class plane): max, sum of exp, the
//             target's score; block sum of the per-point losses in fp64 -> partial[block]
//             (targets equal to ignore_index = -100, NLLLoss's default, do not count; any other
//             out-of-range target makes the loss NaN — loud, PyTorch raises a device assert)
//   kernel 2: fixed-order fp64 sum of the partials, loss = sum / count.
#include "common.cuh"

namespace sonet {

constexpr int SL_THREADS = 256;

__global__ void __launch_bounds__(SL_THREADS)
    seg_loss_partial_kernel(const float* __restrict__ score, const long long* __restrict__ target,
                            int C, int N, double* __restrict__ psum, int* __restrict__ pcnt) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * SL_THREADS + threadIdx.x;
  double loss = 0.0;
  int cnt = 0;
  if (n < N) {
    const long long t = target[static_cast<size_t>(b) * N + n];
    if (t != -100) {
      const float* s = score + static_cast<size_t>(b) * C * N + n;
      float mx = -__int_as_float(0x7f800000);
      for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(s + static_cast<size_t>(c) * N));
      float sum = 0.f;
      for (int c = 0; c < C; ++c) sum += expf(__ldg(s + static_cast<size_t>(c) * N) - mx);
      if (t >= 0 && t < C) {
        const float st = __ldg(s + static_cast<size_t>(t) * N);
        loss = static_cast<double>(-((st - mx) - logf(sum)));
      } else {
        loss = static_cast<double>(__int_as_float(0x7fc00000));
      }
      cnt = 1;
    }
  }
  __shared__ double sl[SL_THREADS / 32];
  __shared__ int sc[SL_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    loss += __shfl_xor_sync(0xffffffffu, loss, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) {
    sl[warp] = loss;
    sc[warp] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    int k = 0;
    for (int w = 0; w < SL_THREADS / 32; ++w) {
      a += sl[w];
      k += sc[w];
    }
    const size_t blk = static_cast<size_t>(b) * gridDim.x + blockIdx.x;
    psum[blk] = a;
    pcnt[blk] = k;
  }
}

__global__ void __launch_bounds__(1024)
    seg_loss_final_kernel(const double* __restrict__ psum, const int* __restrict__ pcnt, int nblk,
                          int size_average, float* __restrict__ loss) {
  // fixed assignment of partials to threads and a fixed-shape tree: bit-reproducible
  double a = 0.0;
  long long k = 0;
  for (int i = threadIdx.x; i < nblk; i += 1024) {
    a += psum[i];
    k += pcnt[i];
  }
  __shared__ double sa[32];
  __shared__ long long sk[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    k += __shfl_xor_sync(0xffffffffu, k, o);
  }
  if (lane == 0) {
    sa[warp] = a;
    sk[warp] = k;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    long long c = 0;
    for (int w = 0; w < 32; ++w) {
      t += sa[w];
      c += sk[w];
    }
    // all targets ignored: PyTorch's mean reduction gives nan (0/0) as well
    *loss = static_cast<float>(size_average ? t / static_cast<double>(c) : t);
  }
}

}  // namespace sonet

extern "C" long long sonet_seg_loss_scratch_bytes(int B, int N) {
  if (B < 0 || N < 0) return -1;
  const long long nblk = static_cast<long long>(B) * ((N + sonet::SL_THREADS - 1) / sonet::SL_THREADS);
  return nblk * (sizeof(double) + sizeof(int)) + 16;
}

extern "C" int sonet_seg_loss_f32(const float* score, const long long* target, int B, int C, int N,
                                  int size_average, void* scratch, float* loss,
                                  sonet_stream_t stream) {
  using namespace sonet;
  SONET_REQUIRE(B >= 1 && C >= 1 && N >= 1, "seg_loss: bad dimension");
  SONET_REQUIRE(B <= 65535, "seg_loss: B=%d exceeds grid limit", B);
  SONET_REQUIRE(score && target && scratch && loss, "seg_loss: null pointer");
  const int nb = (N + SL_THREADS - 1) / SL_THREADS;
  const long long nblk = static_cast<long long>(B) * nb;
  SONET_REQUIRE(nblk < (1LL << 31), "seg_loss: too many blocks");
  double* psum = static_cast<double*>(scratch);
  int* pcnt = reinterpret_cast<int*>(psum + nblk);
  cudaStream_t st = as_stream(stream);
  seg_loss_partial_kernel<<<dim3(nb, B), SL_THREADS, 0, st>>>(score, target, C, N, psum, pcnt);
  const int rc = check_launch("seg_loss_partial");
  if (rc != SONET_OK) return rc;
  seg_loss_final_kernel<<<1, 1024, 0, st>>>(psum, pcnt, static_cast<int>(nblk), size_average, loss);
  return check_launch("seg_loss_final");
}
