// common.cuh — shared helpers for libsonet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sonet_b200.h"

namespace sonet {

// thread-local error string, set by SONET_FAIL / check_launch
void set_error(const char* fmt, ...);

#define SONET_FAIL(code, ...)        \
  do {                               \
    ::sonet::set_error(__VA_ARGS__); \
    return (code);                   \
  } while (0)

#define SONET_REQUIRE(cond, ...)                         \
  do {                                                   \
    if (!(cond)) SONET_FAIL(SONET_ERR_BAD_ARG, __VA_ARGS__); \
  } while (0)

// to be called right after a kernel launch; does not synchronise.
int check_launch(const char* what);

inline cudaStream_t as_stream(sonet_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();           // cached multiprocessor count of the current device
int max_smem_optin();     // cached max opt-in dynamic shared memory per block

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- device helpers ---------------------------------------------------------------------------
// Running per-node maxima of the fused pool (csrc/pointmlp_tc.cu) are order-preserving int keys.
constexpr int POOL_KEY_INIT = static_cast<int>(0x80000000u);  // below the key of every float
// key -> value with the reference's semantics: the max must be > -1000 (index_max.cpp:80-81,103),
// otherwise the node gathers the feature of stacked copy 0 (models/networks.py:185).
__device__ __forceinline__ float pool_key_value(int key, float p0v) {
  const int bits = key ^ ((key >> 31) & 0x7fffffff);
  const float v = __int_as_float(bits);
  return (key != POOL_KEY_INIT && v > -1000.0f) ? v : p0v;
}
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
// same, asking L2 to fetch the whole 256-byte neighbourhood from DRAM (sequential row streams)
__device__ __forceinline__ float4 ldg_stream256_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream_f1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// ---- packed fp32x2 arithmetic (sm_100 FADD2 / FMUL2): IEEE round-to-nearest per half ------------
// Inline PTX with explicit .rn: the CUDA intrinsics __fmul2_rn/__fadd2_rn were observed to be
// contracted into FFMA2 by nvcc 12.9, which changes the rounding of (dx*dx + dy*dy) + dz*dz and
// with it the bit-exact distance parity the arg-min kernels rely on.
__device__ __forceinline__ float2 add2_rn(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(r)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 mul2_rn(float2 a, float2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(r)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
// NOTE: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with explicit .rn and
// under --fmad=false (it also rewrites fma.rn.f32x2(a,b,-0) back into a mul and re-fuses it). Where
// the separate rounding of a product and a sum matters, do the PRODUCT packed and the SUM with the
// scalar __fadd_rn (explicitly rounded scalar ops are never contracted).
__device__ __forceinline__ float2 sqsum3_rn(float2 dx, float2 dy, float2 dz) {
  const float2 xx = mul2_rn(dx, dx), yy = mul2_rn(dy, dy), zz = mul2_rn(dz, dz);
  return make_float2(__fadd_rn(__fadd_rn(xx.x, yy.x), zz.x), __fadd_rn(__fadd_rn(xx.y, yy.y), zz.y));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier / bulk-copy (TMA 1-D) wrappers ------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk async copy global -> shared (TMA unit, SASS UBLKCP); size multiple of 16, 16B aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

}  // namespace sonet
