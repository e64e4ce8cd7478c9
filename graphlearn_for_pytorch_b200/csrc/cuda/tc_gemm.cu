// TMA-fed tcgen05 GEMMs of the GraphSAGE engine (sm_100a): every dense contraction of the training step that
// is not already inside the fused layer-1 kernel runs here instead of in cuBLAS --
//
//   forward  layers 2..L : Z  = act([mean | self] . W^T + b)            A K-major, B K-major, bf16 out
//   backward dA          : dA = dPre . W                                 A K-major, B MN-major, bf16 out
//   backward dW          : dW += dPre^T . A   (split-K over the batch)   A MN-major, B MN-major, fp32 red-add
//
// The extents that depend on the sampled batch (rows of A for forward / dA, the contraction length for dW) are
// read from the sampler's device counters, so the launch is static and CUDA-graph capturable.
//
// One persistent CTA per SM, 6 warps:
//   warp 0 (one lane)  TMA producer: cp.async.bulk.tensor.2d boxes (SWIZZLE_128B) into a 4-stage smem ring
//   warp 1 (one lane)  tcgen05.mma issuer (cta_group::1, M=128, N<=256, K=16 x4 per 64-wide k-block), fp32
//                      accumulators in TMEM, two 256-column stages so tile i+1 accumulates while tile i drains
//   warps 2-5          epilogue: tcgen05.ld -> bias/ReLU -> bf16 -> swizzled smem staging -> TMA store,
//                      or fp32 red.global.add.v4 for the split-K weight gradients
// Up to two independent problems share one launch (dW and dA of a layer both consume dPre): work items of both
// are dealt round-robin to the CTAs.
//
// The reference has no counterpart (its examples call PyG / cuBLAS through autograd).
#include <cuda.h>

#include <cstdio>
#include <cstring>
#include <mutex>

#include "launch_utils.h"
#include "tc_utils.cuh"

namespace glt {

namespace {

using namespace tc;

constexpr int kBM = 128;              // tile rows (UMMA M)
constexpr int kBK = 64;               // k-block: 64 bf16 = 128 B = one swizzle row
constexpr int kStages = 4;
constexpr int kABytes = kBM * 128;    // 16 KB
constexpr int kBBytes = 256 * 128;    // 32 KB (BN <= 256)
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kStoreBytes = kBM * 128;   // one 128 x 64 bf16 staging tile for the TMA store
constexpr int kGemmThreads = 6 * 32;
constexpr int kEpiThreads = 4 * 32;
constexpr size_t kGemmSmem = static_cast<size_t>(kStages) * kStageBytes + 2 * kStoreBytes + 1024 + 256;

struct Derived {   // per-problem quantities every role derives identically from device-side state
  int m_tiles, n_tiles, kb_total, splits, kb_per, items;
};

__device__ __forceinline__ Derived derive(const TcProblem& p, int grid_share) {
  Derived d;
  int dyn = p.dyn ? min(p.dyn[p.dyn_idx], p.dyn_cap) : p.dyn_cap;
  if (dyn < 0) dyn = 0;
  const int m_ext = p.dyn_is_k ? p.m : dyn;
  const int k_ext = p.dyn_is_k ? dyn : p.k;
  d.m_tiles = (m_ext + kBM - 1) / kBM;
  d.n_tiles = p.n / p.bn;
  d.kb_total = (k_ext + kBK - 1) / kBK;
  const int tiles = d.m_tiles * d.n_tiles;
  int splits = 1;
  if (p.epi == 1 && tiles > 0) {
    splits = grid_share / tiles;
    if (splits < 1) splits = 1;
    if (splits > d.kb_total) splits = d.kb_total;
    if (splits < 1) splits = 1;
  }
  d.kb_per = (d.kb_total + splits - 1) / splits;
  d.splits = d.kb_per > 0 ? (d.kb_total + d.kb_per - 1) / d.kb_per : 0;   // no empty split
  d.items = (d.kb_total > 0) ? tiles * d.splits : 0;
  return d;
}

struct Item { int prob, m_tile, n_tile, kb0, kb1; };

__device__ __forceinline__ Item decode(int item, const Derived& d0, const Derived& d1) {
  Item it;
  it.prob = item < d0.items ? 0 : 1;
  const Derived& d = it.prob ? d1 : d0;
  int r = it.prob ? item - d0.items : item;
  const int split = r % d.splits;
  r /= d.splits;
  it.n_tile = r % d.n_tiles;
  it.m_tile = r / d.n_tiles;
  it.kb0 = split * d.kb_per;
  it.kb1 = min(it.kb0 + d.kb_per, d.kb_total);
  return it;
}

__global__ void __launch_bounds__(kGemmThreads, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
          const __grid_constant__ CUtensorMap tmC0, const __grid_constant__ CUtensorMap tmA1,
          const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmC1,
          const TcGemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage0 = smem;                                     // kStages x [A 16 KB | B 32 KB]
  uint8_t* store0 = smem + kStages * kStageBytes;             // 2 x 16 KB staging tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(store0 + 2 * kStoreBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t bar_full0 = smem_u32(bars + 0);               // + 8 * stage
  const uint32_t bar_empty0 = smem_u32(bars + kStages);        // + 8 * stage
  const uint32_t bar_tfull0 = smem_u32(bars + 2 * kStages);    // + 8 * acc stage
  const uint32_t bar_tempty0 = smem_u32(bars + 2 * kStages + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full0 + 8 * s, 1);
      mbar_init(bar_empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull0 + 8 * s, 1);
      mbar_init(bar_tempty0 + 8 * s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&tmA0); prefetch_tmap(&tmB0);
    if (g.p[0].epi == 0) prefetch_tmap(&tmC0);
    if (g.n_prob > 1) {
      prefetch_tmap(&tmA1); prefetch_tmap(&tmB1);
      if (g.p[1].epi == 0) prefetch_tmap(&tmC1);
    }
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
  // everything above (barrier init, tensor-map prefetch, TMEM allocation) overlapped with the tail of the previous
  // kernel; from here on the predecessor's results (counters, operands) are read
  pdl_wait();
  pdl_trigger();

  // identical on every thread: work decomposition from the device counters
  const int share = g.n_prob > 1 ? max(1, static_cast<int>(gridDim.x) / 2) : static_cast<int>(gridDim.x);
  const Derived d0 = derive(g.p[0], share);
  Derived d1; d1.m_tiles = d1.n_tiles = d1.kb_total = d1.splits = d1.kb_per = d1.items = 0;
  if (g.n_prob > 1) d1 = derive(g.p[1], share);
  const int total_items = d0.items + d1.items;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int fill = 0;   // k-blocks produced so far (ring position)
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const Item it = decode(item, d0, d1);
        const TcProblem& p = g.p[it.prob];
        const CUtensorMap* ta = it.prob ? &tmA1 : &tmA0;
        const CUtensorMap* tb = it.prob ? &tmB1 : &tmB0;
        const uint32_t bytes = kABytes + static_cast<uint32_t>(p.bn) * 128u;
        for (int kb = it.kb0; kb < it.kb1; ++kb, ++fill) {
          const int s = fill % kStages;
          const uint32_t ph = (fill / kStages) & 1;
          mbar_wait(bar_empty0 + 8 * s, ph ^ 1);
          const uint32_t full = bar_full0 + 8 * s;
          mbar_expect_tx(full, bytes);
          const uint32_t sa = smem_u32(stage0 + s * kStageBytes);
          const uint32_t sb = sa + kABytes;
          if (!p.a_mn) {
            tma_load_2d(sa, ta, kb * kBK, it.m_tile * kBM, full);              // box {64 k, 128 rows}
          } else {
            tma_load_2d(sa, ta, it.m_tile * kBM, kb * kBK, full);              // box {64 m, 64 k-rows}
            tma_load_2d(sa + 8192, ta, it.m_tile * kBM + 64, kb * kBK, full);
          }
          if (!p.b_mn) {
            tma_load_2d(sb, tb, kb * kBK, it.n_tile * p.bn, full);             // box {64 k, bn rows}
          } else {
            for (int j = 0; j < p.bn / 64; ++j)                                // boxes {64 n, 64 k-rows}
              tma_load_2d(sb + j * 8192, tb, it.n_tile * p.bn + j * 64, kb * kBK, full);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      int drain = 0, acc_it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++acc_it) {
        const Item it = decode(item, d0, d1);
        const TcProblem& p = g.p[it.prob];
        const uint32_t idesc = make_idesc_bf16(kBM, p.bn, p.a_mn, p.b_mn);
        const int acc = acc_it & 1;
        mbar_wait(bar_tempty0 + 8 * acc, ((acc_it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + acc * 256;
        for (int kb = it.kb0; kb < it.kb1; ++kb, ++drain) {
          const int s = drain % kStages;
          mbar_wait(bar_full0 + 8 * s, (drain / kStages) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage0 + s * kStageBytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t da = p.a_mn ? make_sw128_desc_lbo(sa + k4 * 2048, 8192, 1024)
                                       : make_sw128_desc_lbo(sa + k4 * 32, 16, 1024);
            const uint64_t db = p.b_mn ? make_sw128_desc_lbo(sb + k4 * 2048, 8192, 1024)
                                       : make_sw128_desc_lbo(sb + k4 * 32, 16, 1024);
            umma_bf16(tmem_c, da, db, idesc, (kb > it.kb0 || k4 > 0) ? 1u : 0u);
          }
          umma_commit(bar_empty0 + 8 * s);            // smem stage reusable once these MMAs retire
        }
        umma_commit(bar_tfull0 + 8 * acc);            // accumulator complete
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue ------------------------------
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may read
    const int et = threadIdx.x - 64;                   // 0..127 inside the epilogue group
    const bool issuer = (et == 0);
    int acc_it = 0;
    int chunk_ctr = 0;                                 // staging double buffer position
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++acc_it) {
      const Item it = decode(item, d0, d1);
      const TcProblem& p = g.p[it.prob];
      const CUtensorMap* tcm = it.prob ? &tmC1 : &tmC0;
      const int acc = acc_it & 1;
      mbar_wait(bar_tfull0 + 8 * acc, (acc_it >> 1) & 1);
      tc_fence_after();
      const int r = quad * 32 + lane;                  // tile row == TMEM lane
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256;
      if (p.epi == 0) {
        const __nv_bfloat16* bias = reinterpret_cast<const __nv_bfloat16*>(p.bias);
        for (int c0 = 0; c0 < p.bn; c0 += 64, ++chunk_ctr) {
          uint8_t* stg = store0 + (chunk_ctr & 1) * kStoreBytes;
          // the TMA store issued two chunks ago must have finished READING this staging tile
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
                const int col = it.n_tile * p.bn + c0 + h + q * 8;
                float b[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) b[i] = 0.f;
                bf16x8_accum(*reinterpret_cast<const uint4*>(bias + col), b);
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] += b[i];
              }
              if (p.relu) {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fmaxf(x[i], 0.f);
              }
              const int chunk = (h >> 3) + q;          // 16-byte chunk inside the 128-byte staged row
              sts128(srow + ((chunk ^ (r & 7)) << 4), pack_bf16x8(x, 1.f));
            }
          }
          if (c0 + 64 >= p.bn) {                       // last TMEM read of this tile: release the accumulator
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty0 + 8 * acc);
          }
          fence_proxy_async();                         // generic-proxy smem writes -> visible to the TMA engine
          named_bar_sync(1, kEpiThreads);
          if (issuer) {
            tma_store_2d(tcm, it.n_tile * p.bn + c0, it.m_tile * kBM, smem_u32(stg));
            tma_commit_group();
          }
        }
      } else {
        // split-K partial sums: fp32 vector reductions straight into the gradient buffer
        const int row = it.m_tile * kBM + r;
        float* orow = p.out32 + static_cast<int64_t>(row) * p.ld32 + it.n_tile * p.bn;
        for (int c0 = 0; c0 < p.bn; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          if (row < p.m_valid) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + c0 + q * 4),
                           "f"(__uint_as_float(v[q * 4])), "f"(__uint_as_float(v[q * 4 + 1])),
                           "f"(__uint_as_float(v[q * 4 + 2])), "f"(__uint_as_float(v[q * 4 + 3]))
                           : "memory");
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty0 + 8 * acc);
      }
    }
    if (issuer) tma_wait_group<0>();                   // all stores retired before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------------------------------------------
// host side: tensor maps through the driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

}  // namespace

// 2-D bf16 row-major tensor [rows, cols] (cols contiguous, row pitch `ld` elements), box {box_cols, box_rows},
// SWIZZLE_128B (box_cols * 2 bytes must be 128), zero fill out of bounds.
int make_tmap_bf16_2d(void* out_map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                      int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return 1;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(reinterpret_cast<CUtensorMap*>(out_map), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                        const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

int make_tmap_u8_2d(void* out_map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                    int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return 1;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(reinterpret_cast<CUtensorMap*>(out_map), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                        const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

size_t tc_gemm_smem_bytes() { return kGemmSmem; }

void launch_tc_gemm(const TcGemmLaunch& L, int num_sms, cudaStream_t s) {
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(k_tc_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmem));
  });
  // per-device attribute: cheap to repeat, required on every device the kernel runs on
  cudaFuncSetAttribute(k_tc_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmem));
  int grid = num_sms;
  if (L.max_items > 0 && L.max_items < grid) grid = L.max_items;
  const CUtensorMap* m = reinterpret_cast<const CUtensorMap*>(L.maps);
  launch_k(k_tc_gemm, dim3(grid), dim3(kGemmThreads), kGemmSmem, s, m[0], m[1], m[2], m[3], m[4], m[5], L.args);
}

}  // namespace glt
