// pointmlp_tc.cu — the first PointResNet (6|3 -> 64 -> 128 -> 256 -> [64+256] -> 384) of the SO-Net
// encoder as ONE persistent tcgen05 kernel for sm_100a.
//
// Replaces Encoder.first_pointnet = PointResNet.forward (models/layers.py:419-432, wired at
// models/networks.py:82-83,176): four EquivariantLayers (Conv1d k=1 + eval BatchNorm + ReLU,
// layers.py:282-296) and the skip-concat torch.cat((layer0_out, x_tmp)) — 88 % of the forward's
// FLOPs. The reference round-trips every activation through HBM (~7.4 GB at B=64,N=5000); here a
// 128-point tile walks through all four layers inside one SM:
//
//   * points are the MMA M dimension (128 TMEM lanes = 128 point copies), channels are N/K;
//   * layer 0 (K=3|6, degenerate for tensor cores) runs on CUDA cores, one thread per point;
//   * layers 1-3 are tcgen05.mma kind::f16 with the ACTIVATIONS AS THE A OPERAND READ FROM TENSOR
//     MEMORY: the epilogue warps read the fp32 accumulator (tcgen05.ld), add the folded-BN shift,
//     ReLU, split into fp16 hi/lo and write the pair back IN PLACE over the accumulator columns
//     (tcgen05.st) — a 16-channel group of fp32 columns becomes 8 hi + 8 lo packed columns, which
//     is exactly the K-major A layout the next layer's MMA consumes. Activations never touch
//     shared memory or HBM;
//   * fp32 parity (1e-4) on 16-bit tensor cores comes from the 3-product split
//     x*w ~= hi(x)*hi(w) + lo(x)*hi(w) + hi(x)*lo(w) with FP16 halves (hi+lo carry 22 mantissa
//     bits; the same split in bf16 carries 16 and measured 1.1e-4, outside the bar). Weights are
//     pre-scaled per layer by a power of two (undone exactly in the epilogue FMA) so their lo
//     halves stay normal fp16 numbers; activations are clamped to the fp16 range (65504);
//   * weights (B operand, K-major no-swizzle core-matrix images packed once by the host) stream
//     L2 -> shared memory through a 5-slot TMA ring (cp.async.bulk + mbarrier); layer-1 weights
//     stay resident;
//   * MMA shapes are as wide as TMEM allows, because ONE thread issues them and the issue rate is
//     what bounded the first version (480 N=64 MMAs per tile: measured 100 -> 60 cycles per MMA
//     against 32 cycles of tensor work): layer 1 = 12 MMAs of N=128, layer 2 = 24 of N=256,
//     layer 3 = 4 chunks x 60 of N=96 over two accumulator buffers (its epilogue — shift add +
//     coalesced fp32 stores, lane = point — overlaps the next chunk's MMAs): 276 MMAs per tile;
//   * the MMA warp stays converged and `elect.sync` predicates each tcgen05 instruction (issuing
//     from a divergent `if (lane == 0)` makes nvcc wrap every UTCHMMA in an ELECT/BRA loop);
//   * warp roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11
//     epilogue (two warpgroups splitting the columns; warp%4 selects the TMEM lane quarter);
//   * TMEM map (512 columns): [0,64) act0 | [64,320) D2/act2 | [320,448) D1/act1, and after
//     layer 2: [320,416), [416,512) the two D3 buffers.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace sonet {

namespace pm {
constexpr int C0 = 64, C1 = 128, C2 = 256, C3 = 384, K3 = C0 + C2;  // layer widths
constexpr int TILE = 128;                                            // points per tile (MMA M)
constexpr int SLOT_BYTES = 32768, NSLOT = 5;
constexpr int W1_BYTES = 2 * C1 * C0 * 2;                            // hi + lo images, 32 KB
constexpr int L2_STAGES = 4, L2_KT = 32;                             // [256 rows x 32 k] per stage
constexpr int L3_CHUNKS = 4, L3_N = 96, L3_KSTAGES = 5, L3_KT = 64;  // [96 rows x 64 k] per stage
constexpr int L2_STAGE_BYTES = C2 * L2_KT * 2 * 2;                   // 32768
constexpr int L3_STAGE_BYTES = L3_N * L3_KT * 2 * 2;                 // 24576
constexpr int STAGES_PER_TILE = L2_STAGES + L3_CHUNKS * L3_KSTAGES;  // 24
constexpr int BLOB_BYTES =
    W1_BYTES + L2_STAGES * L2_STAGE_BYTES + L3_CHUNKS * L3_KSTAGES * L3_STAGE_BYTES;  // 655360
constexpr int NFP = C0 * 6 + C0 + C1 + C2 + C3 + 4;   // W0 | shift0..3 | 1/wscale1..3 (floats)
constexpr int NUM_THREADS = 384;
constexpr int NBAR = 2 * NSLOT + 5 + 4 + 1;
// shared memory carve-up
constexpr int OFF_W1 = 0;
constexpr int OFF_RING = OFF_W1 + W1_BYTES;
constexpr int OFF_FP = OFF_RING + NSLOT * SLOT_BYTES;
constexpr int OFF_BAR = OFF_FP + NFP * 4;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16;
// TMEM columns
constexpr uint32_t COL_A0 = 0, COL_D2 = 64, COL_D1 = 320, COL_D3 = 320;

__host__ __device__ constexpr int stage_bytes(int s) {
  return s < L2_STAGES ? L2_STAGE_BYTES : L3_STAGE_BYTES;
}
// descriptor: lo word = start address >> 4 | (LBO=128 >> 4) << 16; hi word = SBO >> 4 | version 1.
// Advancing the start address by n bytes is desc + (n >> 4).
__device__ __forceinline__ uint64_t bdesc(uint32_t saddr, uint32_t sbo) {
  const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (8u << 16);
  const uint32_t hi = (sbo >> 4) | (1u << 14);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
}  // namespace pm

// CL = thread-block-cluster size: the CL CTAs of a cluster consume the same weight stream, each
// fetches 1/CL of every stage from L2 and multicasts it into all CL shared memories. Measured: no
// gain at CL=2 and a loss at CL=4 (the stream is not the bound; lockstep coupling costs), so the
// default is CL=1; kept as a tested option (SONET_TC_CLUSTER).
//
// POOL variant (the classifier / auto-encoder path): the input rows are the stacked copies SORTED
// BY NODE (csrc/som_sort.cu) and the layer-3 epilogue, instead of storing the 384 channels of
// every copy (1.47 GB at B=64,N=5000) for a separate index_max launch to re-read, reduces them
// per node right away: lanes of a warp that share a node do one `redux.sync.max` on an
// order-preserving integer key and the group leader one `red.global.max` into pool[b,c,node]
// (models/index_max_ext semantics are restored by pool_finalize_kernel). first_pn_out is never
// written.
struct PoolArgs {
  const int32_t* node_sorted;  // [B,P] node id of each sorted row
  const int32_t* pos0;         // [B] sorted row of stacked copy 0
  int32_t* keys;               // [B,384,M] running max keys (POOL_KEY_INIT when untouched)
  float* p0;                   // [B,384] features of stacked copy 0
  int M;
};

template <int CL, bool POOL>
__global__ void __launch_bounds__(pm::NUM_THREADS, 1)
    pointresnet_tc_kernel(const float* __restrict__ x_in, int Cin, int B, int P,
                          const unsigned char* __restrict__ blob, const float* __restrict__ fparams,
                          float* __restrict__ out, long long* __restrict__ dbg, PoolArgs pool) {
  using namespace pm;
  // optional timeline (debug entry point only): clock64 at phase boundaries of CTA 0's 4th tile
#define PM_TL(role, idx)                                                          \
  do {                                                                            \
    if (dbg != nullptr && blockIdx.x == 0 && t == 3 && lane == 0)                 \
      dbg[(role) * 32 + (idx)] = clock64();                                       \
  } while (0)
  extern __shared__ __align__(1024) unsigned char smem[];
  float* fp = reinterpret_cast<float*>(smem + OFF_FP);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* full = bars;                   // [NSLOT]  TMA -> MMA
  uint64_t* empty = bars + NSLOT;          // [NSLOT]  MMA -> TMA
  uint64_t* b_act0 = bars + 2 * NSLOT;     // epilogue -> MMA
  uint64_t* b_d1 = b_act0 + 1;             // MMA -> epilogue
  uint64_t* b_act1 = b_act0 + 2;
  uint64_t* b_d2 = b_act0 + 3;
  uint64_t* b_act2 = b_act0 + 4;
  uint64_t* d3full = b_act0 + 5;           // [2]
  uint64_t* d3empty = d3full + 2;          // [2]
  uint64_t* w1_full = d3empty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_cloud = (P + TILE - 1) / TILE;
  const int num_tiles = B * tiles_per_cloud;
  // every CTA runs the same number of iterations (a cluster consumes the weight stream in
  // lockstep); iterations past the last tile compute on zeros and store nothing
  const int my_tiles = (num_tiles + gridDim.x - 1) / gridDim.x;
  constexpr uint16_t CL_MASK = static_cast<uint16_t>((1u << CL) - 1);
  const uint32_t cta_rank = (CL > 1) ? tc::cluster_ctarank() : 0;

  for (int i = threadIdx.x; i < NFP; i += NUM_THREADS) fp[i] = fparams[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CL);
    }
    mbar_init(b_act0, 8);
    mbar_init(b_d1, 1);
    mbar_init(b_act1, 8);
    mbar_init(b_d2, 1);
    mbar_init(b_act2, 8);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d3full[i], 1);
      mbar_init(&d3empty[i], 8);
    }
    mbar_init(w1_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tc::tmem_alloc(tmem_ptr, 512);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (CL > 1) tc::cluster_sync_all();   // peers' barriers are initialised before any multicast lands
  tc::fence_after_sync();
  const uint32_t tm = *tmem_ptr;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      mbar_arrive_expect_tx(w1_full, W1_BYTES);
      bulk_g2s(smem + OFF_W1, blob, W1_BYTES, w1_full);
      uint32_t q = 0;
      for (int t = 0; t < my_tiles; ++t) {
        uint32_t off = W1_BYTES;
        for (int s = 0; s < STAGES_PER_TILE; ++s, ++q) {
          const uint32_t slot = q % NSLOT, use = q / NSLOT;
          if (use > 0) tc::mbar_wait_bounded(&empty[slot], (use - 1) & 1, 100 + s);
          const uint32_t bytes = stage_bytes(s);
          mbar_arrive_expect_tx(&full[slot], bytes);   // all CL slices land on this barrier
          if (CL == 1) {
            bulk_g2s(smem + OFF_RING + slot * SLOT_BYTES, blob + off, bytes, &full[slot]);
          } else {
            const uint32_t slice = bytes / CL;
            tc::bulk_g2s_multicast(smem + OFF_RING + slot * SLOT_BYTES + cta_rank * slice,
                                   blob + off + cta_rank * slice, slice, &full[slot], CL_MASK);
          }
          off += bytes;
        }
      }
    }
  } else if (warp == 1) {
    // ============ MMA issuer: the warp stays converged, elect.sync issues from one lane ============
    constexpr uint32_t ID1 = tc::idesc_f16_f32(TILE, C1), ID2 = tc::idesc_f16_f32(TILE, C2),
                       ID3 = tc::idesc_f16_f32(TILE, L3_N);
    const uint32_t w1_addr = smem_u32(smem + OFF_W1);
    const uint32_t ring_addr = smem_u32(smem + OFF_RING);
    tc::mbar_wait_bounded(w1_full, 0, 1);
    uint32_t q = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const uint32_t par = t & 1;
      // ---- layer 1: D1[128 x 128] = act0[128 x 64] * W1^T (one N=128 MMA per product) ----
      tc::mbar_wait_bounded(b_act0, par, 2);
      if (t > 0) {  // D1 overlaps both D3 buffers of the previous tile: wait for their last drain
        tc::mbar_wait_bounded(&d3empty[0], 1, 3);
        tc::mbar_wait_bounded(&d3empty[1], 1, 4);
      }
      tc::fence_after_sync();
      PM_TL(0, 0);
      {
        const uint64_t dh0 = bdesc(w1_addr, 1024), dl0 = bdesc(w1_addr + W1_BYTES / 2, 1024);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t a_hi = tm + COL_A0 + 16 * ks, a_lo = a_hi + 8;
          const uint64_t dh = dh0 + ks * 16, dl = dl0 + ks * 16;   // +256 B per K step
          tc::mma_ts_elect(tm + COL_D1, a_hi, dh, ID1, ks > 0);
          tc::mma_ts_elect(tm + COL_D1, a_lo, dh, ID1, 1);
          tc::mma_ts_elect(tm + COL_D1, a_hi, dl, ID1, 1);
        }
      }
      tc::commit_elect(b_d1);
      PM_TL(0, 1);
      // ---- layer 2: D2[128 x 256] = act1[128 x 128] * W2^T: 4 streamed K slabs of 32, N=256 ----
      tc::mbar_wait_bounded(b_act1, par, 5);
      tc::fence_after_sync();
      PM_TL(0, 2);
#pragma unroll
      for (int kc = 0; kc < L2_STAGES; ++kc, ++q) {
        const uint32_t slot = q % NSLOT;
        tc::mbar_wait_bounded(&full[slot], (q / NSLOT) & 1, 6);
        tc::fence_after_sync();
        const uint32_t sb = ring_addr + slot * SLOT_BYTES;
        const uint64_t dh0 = bdesc(sb, 512), dl0 = bdesc(sb + L2_STAGE_BYTES / 2, 512);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint32_t a_hi = tm + COL_D1 + 16 * (kc * 2 + ks), a_lo = a_hi + 8;
          const uint64_t dh = dh0 + ks * 16, dl = dl0 + ks * 16;
          tc::mma_ts_elect(tm + COL_D2, a_hi, dh, ID2, (kc | ks) != 0);
          tc::mma_ts_elect(tm + COL_D2, a_lo, dh, ID2, 1);
          tc::mma_ts_elect(tm + COL_D2, a_hi, dl, ID2, 1);
        }
        if (CL == 1) tc::commit_elect(&empty[slot]);
        else tc::commit_multicast_elect(&empty[slot], CL_MASK);
      }
      tc::commit_elect(b_d2);
      PM_TL(0, 3);
      // ---- layer 3: D3[128 x 384] = cat(act0, act2)[128 x 320] * W3^T: 4 chunks of N=96 ----
      tc::mbar_wait_bounded(b_act2, par, 7);
      tc::fence_after_sync();
      PM_TL(0, 4);
#pragma unroll 1
      for (int nc = 0; nc < L3_CHUNKS; ++nc) {
        const int buf = nc & 1;
        if (nc >= 2) tc::mbar_wait_bounded(&d3empty[buf], 0, 8);   // drain of chunk nc-2 (this tile)
        tc::fence_after_sync();
        const uint32_t d = tm + COL_D3 + L3_N * buf;
        PM_TL(0, 5 + 2 * nc);
#pragma unroll
        for (int kc = 0; kc < L3_KSTAGES; ++kc, ++q) {
          const uint32_t slot = q % NSLOT;
          tc::mbar_wait_bounded(&full[slot], (q / NSLOT) & 1, 10);
          tc::fence_after_sync();
          const uint32_t sb = ring_addr + slot * SLOT_BYTES;
          const uint64_t dh0 = bdesc(sb, 1024), dl0 = bdesc(sb + L3_STAGE_BYTES / 2, 1024);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const int kg = kc * 4 + ks;  // 16-channel K group: 0-3 act0, 4-19 act2
            const uint32_t a_hi = tm + (kg < 4 ? COL_A0 + 16 * kg : COL_D2 + 16 * (kg - 4));
            const uint32_t a_lo = a_hi + 8;
            const uint64_t dh = dh0 + ks * 16, dl = dl0 + ks * 16;
            tc::mma_ts_elect(d, a_hi, dh, ID3, kg > 0);
            tc::mma_ts_elect(d, a_lo, dh, ID3, 1);
            tc::mma_ts_elect(d, a_hi, dl, ID3, 1);
          }
          if (CL == 1) tc::commit_elect(&empty[slot]);
          else tc::commit_multicast_elect(&empty[slot], CL_MASK);
        }
        tc::commit_elect(&d3full[buf]);
        PM_TL(0, 6 + 2 * nc);
      }
    }
  } else if (warp >= 4) {
    // =========================== epilogue warps ===========================
    const int q4 = warp & 3;            // TMEM lane quarter this warp may access
    const int h = (warp - 4) >> 2;      // column half handled by this warpgroup
    const int r = q4 * 32 + lane;       // point within the tile == TMEM lane
    const uint32_t lane_base = tm + (static_cast<uint32_t>(q4 * 32) << 16);
    const float* W0 = fp;
    const float* sh0 = fp + C0 * 6;
    const float* sh1 = sh0 + C0;
    const float* sh2 = sh1 + C1;
    const float* sh3 = sh2 + C2;
    const float inv1 = sh3[C3], inv2 = sh3[C3 + 1], inv3 = sh3[C3 + 2];  // 1 / weight pre-scale
    // Layer 0 of tile t+1 is software-pipelined into tile t: its 6 inputs are prefetched while the
    // layer-3 MMAs run, and it is computed as soon as the LAST layer-3 chunk of tile t has been read
    // out of TMEM (all MMAs reading act0 are then complete) — before that chunk's pooling/stores —
    // so the MMA warp can start tile t+1 about 2.3k cycles earlier (timeline, profiles/r01_summary.md).
    float xn[6];
    auto prefetch_x = [&](int tn) {
      const int tile_n = blockIdx.x + tn * gridDim.x;
      const int bn = tile_n / tiles_per_cloud;
      const int jn = (tile_n - bn * tiles_per_cloud) * TILE + r;
      const bool vn = (tile_n < num_tiles) && (jn < P);
#pragma unroll
      for (int c = 0; c < 6; ++c)
        xn[c] = (vn && c < Cin) ? __ldg(x_in + (static_cast<size_t>(bn) * Cin + c) * P + jn) : 0.f;
    };
    auto layer0 = [&]() {   // CUDA cores: 32 of the 64 channels per warpgroup, from xn[]
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int ch0 = 32 * h + 16 * g;
        float y[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float* w = W0 + (ch0 + i) * 6;
          float a = sh0[ch0 + i];
#pragma unroll
          for (int c = 0; c < 6; ++c) a = fmaf(w[c], xn[c], a);
          y[i] = fminf(fmaxf(a, 0.f), 65504.f);
        }
        uint32_t wds[16];
        tc::split16_f16(y, wds);
        tc::st16(lane_base + COL_A0 + ch0, wds);
      }
      tc::wait_st();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_act0);
    };
    if (my_tiles > 0) {
      prefetch_x(0);
      layer0();
    }
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = blockIdx.x + t * gridDim.x;
      const int b = tile / tiles_per_cloud;
      const int j = (tile - b * tiles_per_cloud) * TILE + r;
      const bool valid = (tile < num_tiles) && (j < P);
      const uint32_t par = t & 1;
      if (warp == 4) PM_TL(1, 1);

      // ---- layer 1 epilogue: 64 of 128 channels, in place ----
      tc::mbar_wait_bounded(b_d1, par, 20);
      tc::fence_after_sync();
      if (warp == 4) PM_TL(1, 2);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int ch0 = 64 * h + 16 * g;
        uint32_t v[16];
        tc::ld16(lane_base + COL_D1 + ch0, v);
        tc::wait_ld();
        float y[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          y[i] = fminf(fmaxf(fmaf(__uint_as_float(v[i]), inv1, sh1[ch0 + i]), 0.f), 65504.f);
        tc::split16_f16(y, v);
        tc::st16(lane_base + COL_D1 + ch0, v);
      }
      tc::wait_st();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_act1);
      if (warp == 4) PM_TL(1, 3);

      // ---- layer 2 epilogue: 128 of 256 channels, in place ----
      tc::mbar_wait_bounded(b_d2, par, 21);
      tc::fence_after_sync();
      if (warp == 4) PM_TL(1, 4);
#pragma unroll 2
      for (int g = 0; g < 8; ++g) {
        const int ch0 = 128 * h + 16 * g;
        uint32_t v[16];
        tc::ld16(lane_base + COL_D2 + ch0, v);
        tc::wait_ld();
        float y[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          y[i] = fminf(fmaxf(fmaf(__uint_as_float(v[i]), inv2, sh2[ch0 + i]), 0.f), 65504.f);
        tc::split16_f16(y, v);
        tc::st16(lane_base + COL_D2 + ch0, v);
      }
      tc::wait_st();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_act2);
      if (warp == 4) PM_TL(1, 5);
      if (t + 1 < my_tiles) prefetch_x(t + 1);   // in flight during the layer-3 MMAs

      // ---- layer 3 epilogue: four chunks of 96 channels, 48 per warpgroup; bare layer (no ReLU) ----
      float* orow = POOL ? nullptr : out + (static_cast<size_t>(b) * C3) * P + j;
      // POOL: lanes sharing a node form a group (rows are node-sorted: 1-2 groups per warp)
      // (full-mask redux + a warp-uniform mode: a redux over match_any masks makes nvcc emit a
      // per-group uniformisation loop, measured 2x slower than not fusing at all)
      constexpr int POOL_KEY_MIN = static_cast<int>(0x80000000u);
      int nd = -1, pmode = -1, laneA = 0, laneB = 0;
      bool inA = false, inB = false, is_p0 = false, any_p0 = false;
      int32_t *krow = nullptr, *krowA = nullptr, *krowB = nullptr;
      if (POOL) {
        if (valid) nd = __ldg(pool.node_sorted + static_cast<size_t>(b) * P + j);
        is_p0 = valid && (j == __ldg(pool.pos0 + b));
        any_p0 = __any_sync(0xffffffffu, is_p0);
        int32_t* kb = pool.keys + (static_cast<size_t>(b) * C3) * pool.M;
        krow = kb + max(nd, 0);
        const unsigned vmask = __ballot_sync(0xffffffffu, nd >= 0);
        if (vmask != 0) {
          laneA = __ffs(vmask) - 1;
          const int nodeA = __shfl_sync(0xffffffffu, nd, laneA);
          inA = (nd == nodeA);
          krowA = kb + nodeA;
          const unsigned rest = vmask & ~__ballot_sync(0xffffffffu, inA);
          if (rest == 0) {
            pmode = 0;
          } else {
            laneB = __ffs(rest) - 1;
            const int nodeB = __shfl_sync(0xffffffffu, nd, laneB);
            inB = (nd == nodeB);
            krowB = kb + nodeB;
            pmode = ((rest & ~__ballot_sync(0xffffffffu, inB)) == 0) ? 1 : 2;
          }
        }
      }
      for (int nc = 0; nc < L3_CHUNKS; ++nc) {
        const int buf = nc & 1;
        tc::mbar_wait_bounded(&d3full[buf], (nc >> 1) & 1, 22);   // two uses per tile: parity = use&1
        tc::fence_after_sync();
        if (warp == 4) PM_TL(1, 6 + 2 * nc);
        uint32_t v0[16], v1[16], v2[16];
        const uint32_t cb = lane_base + COL_D3 + L3_N * buf + 48 * h;
        tc::ld16(cb, v0);
        tc::ld16(cb + 16, v1);
        tc::ld16(cb + 32, v2);
        tc::wait_ld();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d3empty[buf]);
        if (nc == L3_CHUNKS - 1 && t + 1 < my_tiles) {
          if (warp == 4) PM_TL(1, 0);
          layer0();                                // next tile's layer 0 first, then this chunk's work
        }
        if (POOL) {
          const int co0 = L3_N * nc + 48 * h;
          // value -> order-preserving int key; lanes outside `in` contribute the identity
          auto keyof = [&](uint32_t raw, int co, float& val) {
            val = fmaf(__uint_as_float(raw), inv3, sh3[co]);
            const int bits = __float_as_int(val);
            return bits ^ ((bits >> 31) & 0x7fffffff);
          };
          // Per-node max of 16 channels over the warp's 32 rows as a shuffle transpose-reduce:
          // recursive halving (xor 16, 8, 4, 2, then 1) leaves the max of channel
          // c(lane) = lane bits 4..1 (bit 4 = MSB) in every lane pair after 16 shuffles — one
          // per channel — and the even lanes issue ONE 16-lane red. (48 redux.sync per chunk
          // serialise on two uniform registers: measured 4.7-5.8k cycles per chunk, longer than the
          // chunk's MMAs, and the busy epilogue warps starved the MMA warp of issue slots.)
          auto tmax16 = [&](const int (&k16)[16]) {
            int a8[8], b4[4], c2[2];
            {
              const bool hi = (lane & 16) != 0;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int keep = hi ? k16[8 + i] : k16[i], send = hi ? k16[i] : k16[8 + i];
                a8[i] = max(keep, __shfl_xor_sync(0xffffffffu, send, 16));
              }
            }
            {
              const bool hi = (lane & 8) != 0;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int keep = hi ? a8[4 + i] : a8[i], send = hi ? a8[i] : a8[4 + i];
                b4[i] = max(keep, __shfl_xor_sync(0xffffffffu, send, 8));
              }
            }
            {
              const bool hi = (lane & 4) != 0;
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const int keep = hi ? b4[2 + i] : b4[i], send = hi ? b4[i] : b4[2 + i];
                c2[i] = max(keep, __shfl_xor_sync(0xffffffffu, send, 4));
              }
            }
            const bool hi = (lane & 2) != 0;
            const int keep = hi ? c2[1] : c2[0], send = hi ? c2[0] : c2[1];
            const int d1 = max(keep, __shfl_xor_sync(0xffffffffu, send, 2));
            return max(d1, __shfl_xor_sync(0xffffffffu, d1, 1));
          };
          const int my_chan = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 +
                              ((lane >> 1) & 1);
          auto pool_group = [&](const uint32_t (&v)[16], int cbase) {
            if (pmode == 0 || pmode == 1) {
              int keys[16], ka[16];
              float vals;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                keys[i] = keyof(v[i], cbase + i, vals);
                ka[i] = inA ? keys[i] : POOL_KEY_MIN;
              }
              int mine = tmax16(ka);
              if (pmode == 1) {        // warp-uniform and rare: a node boundary inside the warp
                int kb[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) kb[i] = inB ? keys[i] : POOL_KEY_MIN;
                const int mb = tmax16(kb);
                mine = (lane & 1) ? mb : mine;     // odd lanes carry the second node's maxima
              }
              if (!(lane & 1) || pmode == 1) {
                int32_t* dst = ((lane & 1) ? krowB : krowA) +
                               static_cast<size_t>(cbase + my_chan) * pool.M;
                atomicMax(dst, mine);
              }
            } else if (pmode == 2) {   // three or more nodes in one warp (tiny nodes): per lane
              float vals;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int key = keyof(v[i], cbase + i, vals);
                if (nd >= 0) atomicMax(krow + static_cast<size_t>(cbase + i) * pool.M, key);
              }
            }
            if (any_p0) {              // warp-uniform and rare: one warp per cloud and column half
              if (is_p0) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  pool.p0[static_cast<size_t>(b) * C3 + cbase + i] =
                      fmaf(__uint_as_float(v[i]), inv3, sh3[cbase + i]);
              }
            }
          };
          pool_group(v0, co0);
          pool_group(v1, co0 + 16);
          pool_group(v2, co0 + 32);
        } else if (valid) {
          const int co0 = L3_N * nc + 48 * h;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            orow[static_cast<size_t>(co0 + i) * P] = fmaf(__uint_as_float(v0[i]), inv3, sh3[co0 + i]);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            orow[static_cast<size_t>(co0 + 16 + i) * P] =
                fmaf(__uint_as_float(v1[i]), inv3, sh3[co0 + 16 + i]);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            orow[static_cast<size_t>(co0 + 32 + i) * P] =
                fmaf(__uint_as_float(v2[i]), inv3, sh3[co0 + 32 + i]);
        }
        if (warp == 4) PM_TL(1, 7 + 2 * nc);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (CL > 1) tc::cluster_sync_all();   // no CTA leaves while peers may still signal its barriers
  if (warp == 2) tc::tmem_dealloc(tm, 512);
#undef PM_TL
}

// ---- host-side packing -------------------------------------------------------------------------------
// fp16 hi/lo images of scale * W[rows r0..r0+nr) x [k0..k0+kt) (row stride ld), K-major no-swizzle:
// element (r,k) at (r/8)*SBO + (k/8)*128 + (r%8)*16 + (k%8)*2 with SBO = kt*16.
static void pack_tile(const float* W, float scale, int ld, int r0, int nr, int k0, int kt,
                      unsigned char* hi, unsigned char* lo) {
  const uint32_t sbo = static_cast<uint32_t>(kt) * 16;
  for (int r = 0; r < nr; ++r)
    for (int k = 0; k < kt; ++k) {
      const float w = W[static_cast<size_t>(r0 + r) * ld + k0 + k] * scale;  // exact (power of 2)
      const __half h = __float2half_rn(w);
      const __half l = __float2half_rn(w - __half2float(h));
      const uint32_t off = (r >> 3) * sbo + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2;
      std::memcpy(hi + off, &h, 2);
      std::memcpy(lo + off, &l, 2);
    }
}
// power of two that brings max|W| into [256, 512): keeps the lo parts normal fp16 numbers and the
// hi parts far from the 65504 overflow; undone exactly by the epilogue's multiply.
static float pow2_scale(const float* W, size_t n) {
  float m = 0.f;
  for (size_t i = 0; i < n; ++i) m = std::max(m, std::fabs(W[i]));
  if (!(m > 0.f) || !std::isfinite(m)) return 1.f;
  int e;
  std::frexp(m, &e);  // m = f * 2^e, f in [0.5, 1)
  return std::ldexp(1.f, 9 - e);
}

}  // namespace sonet

extern "C" int sonet_pointresnet_tc_blob_bytes(void) { return sonet::pm::BLOB_BYTES; }
extern "C" int sonet_pointresnet_tc_fparam_count(void) { return sonet::pm::NFP; }

extern "C" int sonet_pointresnet_tc_pack(const float* W0, int Cin, const float* W1, const float* W2,
                                         const float* W3, const float* shift0, const float* shift1,
                                         const float* shift2, const float* shift3, void* blob_host,
                                         float* fparams_host) {
  using namespace sonet;
  using namespace sonet::pm;
  SONET_REQUIRE(Cin >= 1 && Cin <= 6, "pointresnet_tc_pack: Cin=%d out of range [1,6]", Cin);
  SONET_REQUIRE(W0 && W1 && W2 && W3 && shift0 && shift1 && shift2 && shift3 && blob_host &&
                    fparams_host,
                "pointresnet_tc_pack: null pointer");
  unsigned char* blob = static_cast<unsigned char*>(blob_host);
  std::memset(blob, 0, BLOB_BYTES);
  const float s1 = pow2_scale(W1, static_cast<size_t>(C1) * C0),
              s2 = pow2_scale(W2, static_cast<size_t>(C2) * C1),
              s3 = pow2_scale(W3, static_cast<size_t>(C3) * K3);
  // layer 1, resident: [128 rows x 64 k]
  pack_tile(W1, s1, C0, 0, C1, 0, C0, blob, blob + W1_BYTES / 2);
  size_t off = W1_BYTES;
  // layer 2: four K slabs [256 rows x 32 k]
  for (int kc = 0; kc < L2_STAGES; ++kc) {
    pack_tile(W2, s2, C1, 0, C2, L2_KT * kc, L2_KT, blob + off, blob + off + L2_STAGE_BYTES / 2);
    off += L2_STAGE_BYTES;
  }
  // layer 3: four 96-row chunks x five K slabs [96 rows x 64 k]
  for (int nc = 0; nc < L3_CHUNKS; ++nc)
    for (int kc = 0; kc < L3_KSTAGES; ++kc) {
      pack_tile(W3, s3, K3, L3_N * nc, L3_N, L3_KT * kc, L3_KT, blob + off,
                blob + off + L3_STAGE_BYTES / 2);
      off += L3_STAGE_BYTES;
    }
  if (off != static_cast<size_t>(BLOB_BYTES)) SONET_FAIL(SONET_ERR_BAD_ARG, "pack: size mismatch");
  float* f = fparams_host;
  for (int c = 0; c < C0; ++c)
    for (int i = 0; i < 6; ++i) f[c * 6 + i] = (i < Cin) ? W0[c * Cin + i] : 0.f;
  std::memcpy(f + C0 * 6, shift0, C0 * 4);
  std::memcpy(f + C0 * 6 + C0, shift1, C1 * 4);
  std::memcpy(f + C0 * 6 + C0 + C1, shift2, C2 * 4);
  std::memcpy(f + C0 * 6 + C0 + C1 + C2, shift3, C3 * 4);
  float* inv = f + C0 * 6 + C0 + C1 + C2 + C3;
  inv[0] = 1.f / s1;
  inv[1] = 1.f / s2;
  inv[2] = 1.f / s3;
  inv[3] = 0.f;
  return SONET_OK;
}

static int launch_pointresnet_tc(const float* x, int Cin, int B, int P, const void* blob,
                                 const float* fparams, float* out, long long* dbg,
                                 const sonet::PoolArgs* pool, sonet_stream_t stream) {
  using namespace sonet;
  using namespace sonet::pm;
  SONET_REQUIRE(B >= 0 && P >= 0, "pointresnet_tc: negative dimension");
  SONET_REQUIRE(Cin >= 1 && Cin <= 6, "pointresnet_tc: Cin=%d out of range [1,6]", Cin);
  if (B == 0 || P == 0) return SONET_OK;
  SONET_REQUIRE(x && blob && fparams && (out || pool), "pointresnet_tc: null pointer");
  SONET_REQUIRE(!pool || (pool->node_sorted && pool->pos0 && pool->keys && pool->p0 && pool->M >= 1),
                "pointresnet_tc: incomplete pool arguments");
  SONET_REQUIRE(aligned16(blob), "pointresnet_tc: weight blob must be 16-byte aligned");
  const long long tiles = static_cast<long long>(B) * ((P + TILE - 1) / TILE);
  SONET_REQUIRE(tiles < (1LL << 31), "pointresnet_tc: too many tiles");
  SONET_REQUIRE(SMEM_BYTES <= max_smem_optin(), "pointresnet_tc: needs %d B of shared memory",
                SMEM_BYTES);
  // cluster size: 1 by default; SONET_TC_CLUSTER=2|4 enables the multicast weight stream
  static int cl_env = -1;
  if (cl_env < 0) {
    const char* e = getenv("SONET_TC_CLUSTER");
    cl_env = e ? atoi(e) : 1;
    if (cl_env != 1 && cl_env != 2 && cl_env != 4) cl_env = 1;
  }
  int cl = pool ? 1 : cl_env;
  const int sms = sm_count();
  while (cl > 1 && (tiles < cl || sms % cl != 0)) cl >>= 1;
  int grid = static_cast<int>(std::min<long long>(tiles, sms));
  grid -= grid % cl;
  auto kern = pool ? pointresnet_tc_kernel<1, true>
                   : (cl == 4 ? pointresnet_tc_kernel<4, false>
                              : (cl == 2 ? pointresnet_tc_kernel<2, false>
                                         : pointresnet_tc_kernel<1, false>));
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PoolArgs pa = pool ? *pool : PoolArgs{nullptr, nullptr, nullptr, nullptr, 0};
  cudaLaunchKernelEx(&cfg, kern, x, Cin, B, P, static_cast<const unsigned char*>(blob), fparams, out,
                     dbg, pa);
  return check_launch("pointresnet_tc");
}

extern "C" int sonet_pointresnet_tc_forward(const float* x, int Cin, int B, int P, const void* blob,
                                            const float* fparams, float* out,
                                            sonet_stream_t stream) {
  return launch_pointresnet_tc(x, Cin, B, P, blob, fparams, out, nullptr, nullptr, stream);
}

extern "C" int sonet_pointresnet_tc_pool_forward(const float* x_sorted, int Cin, int B, int P,
                                                 const void* blob, const float* fparams,
                                                 const int32_t* node_sorted, const int32_t* pos0,
                                                 int M, int32_t* pool_keys, float* p0,
                                                 sonet_stream_t stream) {
  sonet::PoolArgs pa{node_sorted, pos0, pool_keys, p0, M};
  return launch_pointresnet_tc(x_sorted, Cin, B, P, blob, fparams, nullptr, nullptr, &pa, stream);
}

extern "C" int sonet_debug_pointresnet_tc_pool_timeline(const float* x_sorted, int Cin, int B, int P,
                                                        const void* blob, const float* fparams,
                                                        const int32_t* node_sorted,
                                                        const int32_t* pos0, int M,
                                                        int32_t* pool_keys, float* p0,
                                                        long long* timeline64,
                                                        sonet_stream_t stream) {
  sonet::PoolArgs pa{node_sorted, pos0, pool_keys, p0, M};
  return launch_pointresnet_tc(x_sorted, Cin, B, P, blob, fparams, nullptr, timeline64, &pa, stream);
}

extern "C" int sonet_debug_pointresnet_tc_timeline(const float* x, int Cin, int B, int P,
                                                   const void* blob, const float* fparams,
                                                   float* out, long long* timeline64,
                                                   sonet_stream_t stream) {
  return launch_pointresnet_tc(x, Cin, B, P, blob, fparams, out, timeline64, nullptr, stream);
}
