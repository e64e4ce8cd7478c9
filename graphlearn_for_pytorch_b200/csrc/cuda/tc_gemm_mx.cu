// Block-scaled MXFP8 GEMM on tcgen05 (sm_100a):  Z = act(dequant(A) . dequant(W)^T + b)
//
//   A [M, K] e4m3, W [N, K] e4m3 (K-major, 128 elements = one SWIZZLE_128B row), one UE8M0 scale per 32 elements of K
//   tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale  (M=128, N<=256, K=32 per instruction), fp32 accumulators in
//   TMEM, scale factors staged global -> smem (cp.async.bulk) -> TMEM (tcgen05.cp 32x128b.warpx4), bf16 output
//   through a TMA store.
//
// This is the OCP-microscaling counterpart of tc_gemm.cu for inference / serving style forward passes over MXFP8
// activations and weights (data/quantize.py produces both the element bytes and the packed scale blocks).  Scale
// factors travel in the tensor core's native block layout: for every (128 rows x 128 K-elements) block 512 bytes,
// byte (r % 32) * 16 + (r / 32) * 4 + k holds the scale of row r, K-group k (CUTLASS Sm1xxBlockScaledBasicChunk);
// after tcgen05.cp the 16 bytes of a 32-row group sit in four 32-bit TMEM columns, and the byte of K-group k is
// selected by the sf_id field of the instruction descriptor.
//
// Roles (6 warps, one persistent CTA per SM) as in tc_gemm.cu: TMA producer, MMA issuer, 4 epilogue warps.
// No reference counterpart.
#include <cuda.h>

#include "launch_utils.h"
#include "tc_utils.cuh"

namespace glt {

namespace {

using namespace tc;

constexpr int kBM = 128;
constexpr int kBKe = 128;                 // K elements per k-block (128 bytes of e4m3)
constexpr int kStages = 3;
constexpr int kABytes = kBM * 128;        // 16 KB
constexpr int kBBytes = 256 * 128;        // 32 KB
constexpr int kSfBytes = 2048;            // SFA 512 B + SFB up to 1024 B (+pad)
constexpr int kStageBytes = kABytes + kBBytes + kSfBytes;
constexpr int kStoreBytes = kBM * 128;
constexpr int kThreads = 6 * 32;
constexpr int kEpiThreads = 4 * 32;
constexpr size_t kSmem = static_cast<size_t>(kStages) * kStageBytes + 2 * kStoreBytes + 1024 + 256;
constexpr uint32_t kSfCol0 = 256;         // TMEM columns [256, 256 + 16 * kStages): scale factors; [0, 256): accumulator

// shared-memory descriptor of a 32-row x 16-byte scale-factor slab (no swizzle: 8-row core matrices 128 B apart)
__device__ __forceinline__ uint64_t make_sf_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;            // LBO (unused for a 16-byte wide slab)
  d |= static_cast<uint64_t>(128 >> 4) << 32;     // SBO: next 8-row core matrix
  d |= static_cast<uint64_t>(1) << 46;            // descriptor version
  return d;                                       // layout type 0 = SWIZZLE_NONE
}

__device__ __forceinline__ void tmem_cp_sf(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

__device__ __forceinline__ void umma_mxf8(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t tmem_sfa, uint32_t tmem_sfb, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%4], [%5], p;\n"
      "}\n" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(tmem_sfa), "r"(tmem_sfb), "r"(accumulate)
      : "memory");
}

// instruction descriptor (cute::UMMA::InstrDescriptorBlockScaled): e4m3 x e4m3, UE8M0 scales, K-major operands
__device__ __forceinline__ uint32_t make_idesc_mx(int m, int n, uint32_t a_sf_id, uint32_t b_sf_id) {
  return ((b_sf_id & 3u) << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (1u << 23) |
         (static_cast<uint32_t>(m >> 4) << 24) | ((a_sf_id & 3u) << 29);
}

__global__ void __launch_bounds__(kThreads, 1)
k_tc_gemm_mx(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmC, const TcMxArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage0 = smem;
  uint8_t* store0 = smem + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(store0 + 2 * kStoreBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t bar_full0 = smem_u32(bars + 0);
  const uint32_t bar_empty0 = smem_u32(bars + kStages);
  const uint32_t bar_tfull = smem_u32(bars + 2 * kStages);
  const uint32_t bar_tempty = smem_u32(bars + 2 * kStages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full0 + 8 * s, 1);
      mbar_init(bar_empty0 + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmC);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  int m_ext = g.dyn ? min(g.dyn[g.dyn_idx], g.m) : g.m;
  if (m_ext < 0) m_ext = 0;
  const int m_tiles = (m_ext + kBM - 1) / kBM;
  const int n_tiles = g.n / g.bn;
  const int kb_total = g.k / kBKe;
  const int total_items = m_tiles * n_tiles;
  const int nb128 = g.bn / 128;                 // 128-row scale blocks of B per tile (bn = 128 or 256)

  if (warp == 0) {
    if (lane == 0) {
      int fill = 0;
      const uint32_t bytes = kABytes + static_cast<uint32_t>(g.bn) * 128u + 512u + 512u * nb128;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int m_tile = item / n_tiles, n_tile = item % n_tiles;
        for (int kb = 0; kb < kb_total; ++kb, ++fill) {
          const int s = fill % kStages;
          mbar_wait(bar_empty0 + 8 * s, ((fill / kStages) & 1) ^ 1);
          const uint32_t full = bar_full0 + 8 * s;
          mbar_expect_tx(full, bytes);
          const uint32_t sa = smem_u32(stage0 + s * kStageBytes);
          const uint32_t sb = sa + kABytes, ssf = sb + kBBytes;
          tma_load_2d(sa, &tmA, kb * kBKe, m_tile * kBM, full);                  // box {128 B of K, 128 rows}
          tma_load_2d(sb, &tmB, kb * kBKe, n_tile * g.bn, full);                 // box {128 B of K, bn rows}
          bulk_g2s(ssf, reinterpret_cast<const uint8_t*>(g.sfa) + (static_cast<size_t>(m_tile) * kb_total + kb) * 512,
                   512, full);
          for (int j = 0; j < nb128; ++j)
            bulk_g2s(ssf + 512 + j * 512,
                     reinterpret_cast<const uint8_t*>(g.sfb) +
                         (static_cast<size_t>(n_tile * nb128 + j) * kb_total + kb) * 512,
                     512, full);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int drain = 0, acc_it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++acc_it) {
        mbar_wait(bar_tempty, (acc_it & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < kb_total; ++kb, ++drain) {
          const int s = drain % kStages;
          mbar_wait(bar_full0 + 8 * s, (drain / kStages) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage0 + s * kStageBytes);
          const uint32_t sb = sa + kABytes, ssf = sb + kBBytes;
          const uint32_t tsfa = tmem_base + kSfCol0 + s * 16;       // 4 columns: rows 0..127 of A
          const uint32_t tsfb = tsfa + 4;                           // 4 (bn = 128) or 8 (bn = 256) columns
          tmem_cp_sf(tsfa, make_sf_desc(ssf));
          for (int j = 0; j < nb128; ++j) tmem_cp_sf(tsfb + 4 * j, make_sf_desc(ssf + 512 + j * 512));
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {      // K-group k4: scale byte k4 of every 32-bit scale column
            const uint32_t idesc = make_idesc_mx(kBM, g.bn, k4, k4);
            umma_mxf8(tmem_base, make_sw128_desc_lbo(sa + k4 * 32, 16, 1024), make_sw128_desc_lbo(sb + k4 * 32, 16, 1024),
                      idesc, tsfa | (static_cast<uint32_t>(k4) << 30), tsfb | (static_cast<uint32_t>(k4) << 30),
                      (kb > 0 || k4 > 0) ? 1u : 0u);
          }
          umma_commit(bar_empty0 + 8 * s);
        }
        umma_commit(bar_tfull);
      }
    }
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int et = threadIdx.x - 64;
    const bool issuer = (et == 0);
    int acc_it = 0, chunk_ctr = 0;
    const __nv_bfloat16* bias = reinterpret_cast<const __nv_bfloat16*>(g.bias);
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++acc_it) {
      const int m_tile = item / n_tiles, n_tile = item % n_tiles;
      mbar_wait(bar_tfull, acc_it & 1);
      tc_fence_after();
      const int r = quad * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      for (int c0 = 0; c0 < g.bn; c0 += 64, ++chunk_ctr) {
        uint8_t* stg = store0 + (chunk_ctr & 1) * kStoreBytes;
        if (issuer) tma_wait_group_read<1>();
        named_bar_sync(1, kEpiThreads);
        const uint32_t srow = smem_u32(stg) + r * 128;
#pragma unroll
        for (int h = 0; h < 64; h += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0 + h, v);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(v[q * 8 + i]);
            if (bias) {
              float b[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) b[i] = 0.f;
              bf16x8_accum(*reinterpret_cast<const uint4*>(bias + n_tile * g.bn + c0 + h + q * 8), b);
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] += b[i];
            }
            if (g.relu) {
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = fmaxf(x[i], 0.f);
            }
            const int chunk = (h >> 3) + q;
            sts128(srow + ((chunk ^ (r & 7)) << 4), pack_bf16x8(x, 1.f));
          }
        }
        if (c0 + 64 >= g.bn) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty);
        }
        fence_proxy_async();
        named_bar_sync(1, kEpiThreads);
        if (issuer) {
          tma_store_2d(&tmC, n_tile * g.bn + c0, m_tile * kBM, smem_u32(stg));
          tma_commit_group();
        }
      }
    }
    if (issuer) tma_wait_group<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace

void launch_tc_gemm_mx(const void* maps3, const TcMxArgs& a, int num_sms, cudaStream_t s) {
  cudaFuncSetAttribute(k_tc_gemm_mx, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmem));
  const int items = ((a.m + kBM - 1) / kBM) * (a.n / a.bn);
  const int grid = items < num_sms ? (items > 0 ? items : 1) : num_sms;
  const CUtensorMap* m = reinterpret_cast<const CUtensorMap*>(maps3);
  launch_k(k_tc_gemm_mx, dim3(grid), dim3(kThreads), kSmem, s, m[0], m[1], m[2], a);
}

}  // namespace glt
