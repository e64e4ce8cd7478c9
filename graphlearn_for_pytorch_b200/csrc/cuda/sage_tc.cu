// Fused  gather -> neighbour-mean -> [mean | self] x W -> bias/ReLU  for sm_100a.
//
// This is the headline kernel of the engine: the A operand of the first-layer
// GraphSAGE GEMM is never materialised in HBM.  Producer warps pull feature rows
// straight out of the range-partitioned feature table -- local HBM *or a peer
// GPU's HBM over NVLink* (plain ld.global.nc on IPC-mapped addresses) -- reduce
// them to the neighbour mean in registers and write bf16 tiles into
// SWIZZLE_128B K-major shared memory; one elected thread issues tcgen05.mma
// (cta_group::1, M=128, N<=256, K=16) against the weight image that a single
// cp.async.bulk (TMA bulk copy) parked in shared memory; fp32 accumulators live
// in TMEM (two 256-column stages) and four epilogue warps drain them with
// tcgen05.ld, add bias, apply ReLU and store bf16.
//
// It replaces, in one launch, the reference's GatherTensorKernel
// (csrc/cuda/unified_tensor.cu:47-81) + PyG scatter-mean + two cuBLAS GEMMs +
// bias/ReLU elementwise kernels, and keeps the row count on the device.
//
// Roles (672 threads, one persistent CTA per SM):
//   warps 0-15  producers   (8-lane group per target row, 2 rows per group per tile)
//   warp  16    TMEM alloc, weight bulk-load, MMA issue (one elected lane)
//   warps 17-20 epilogue    (TMEM lane quadrant = warp_id % 4)
#include <cstdlib>

#include "launch_utils.h"
#include "tc_utils.cuh"

namespace glt {

namespace {

using namespace tc;

constexpr int kTileM = 128;
constexpr int kProducerWarps = 16;
constexpr int kMmaWarp = 16;
constexpr int kThreads = 21 * 32;
constexpr int kChunkBytes = 128;                    // 64 bf16 = one SWIZZLE_128B atom row
constexpr int kAChunkTile = kTileM * kChunkBytes;   // 16 KB per K-chunk of an A tile

struct HopLoc2 { int hop; int row; };
__device__ __forceinline__ HopLoc2 locate2(const int32_t* cum, int n_hops, int t) {
  HopLoc2 l; l.hop = 0; l.row = t;
#pragma unroll 1
  for (int h = 0; h < n_hops; ++h) {
    const int b = cum[h], e = cum[h + 1];
    if (t >= b && t < e) { l.hop = h; l.row = t - b; break; }
  }
  return l;
}

__device__ __forceinline__ const uint8_t* src_row2(const SageAggArgs& a, int s) {
  if (a.src_local) return reinterpret_cast<const uint8_t*>(a.src_local) + static_cast<int64_t>(s) * a.d * 2;
  return row_ptr(a.feat, a.nodes[s]);
}

// NC = d / 64 (K chunks of the mean half; the self half has NC more)
template <int NC>
__global__ void __launch_bounds__(kThreads, 1) k_sage_fused(SageFusedArgs f) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-B alignment required by SWIZZLE_128B
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int N = f.n_out;
  constexpr int NKC = 2 * NC;                      // K chunks of [mean | self]
  uint8_t* smem_w = smem;                          // NKC x [N x 128 B]
  uint8_t* smem_a = smem_w + NKC * N * kChunkBytes;  // NKC x [128 x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + NKC * kAChunkTile);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t bar_w = smem_u32(bars + 0);
  const uint32_t bar_a_full = smem_u32(bars + 1);
  const uint32_t bar_a_empty = smem_u32(bars + 2);
  const uint32_t bar_t_full0 = smem_u32(bars + 3);   // +8 bytes for stage 1
  const uint32_t bar_t_empty0 = smem_u32(bars + 5);  // +8 bytes for stage 1

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const SageAggArgs& a = f.agg;
  const int T = min(a.cum[a.n_hops_targets], a.cap_targets);
  const int n_tiles = (T + kTileM - 1) / kTileM;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_a_full, kProducerWarps);
    mbar_init(bar_a_empty, 1);
    mbar_init(bar_t_full0, 1);
    mbar_init(bar_t_full0 + 8, 1);
    mbar_init(bar_t_empty0, 4);
    mbar_init(bar_t_empty0 + 8, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kProducerWarps) {
    // ------------------------------ producers ------------------------------
    // 64 eight-lane groups; each owns rows {g, g+64} of the 128-row tile.  The pointer
    // chase (deg -> ELL -> node id -> owner shard) of both rows is issued up front so
    // the four dependent-load chains overlap; feature rows then stream in batches of
    // four neighbours x NC 16-byte vectors per lane.
    const int gl = lane & 7;
    const int gw = lane >> 3;
    const unsigned gmask = 0xFFu << (gw * 8);
    const int group = warp * 4 + gw;  // 0..63
    constexpr int RPG = kTileM / (kProducerWarps * 4);  // rows per group per tile (2)
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      int dg[RPG];
      const int32_t* ell[RPG];
      const uint8_t* p_lo[RPG];
      const uint8_t* p_hi[RPG];
      const uint8_t* p_self[RPG];
#pragma unroll
      for (int q = 0; q < RPG; ++q) {
        const int t = tile * kTileM + group + q * (kTileM / RPG);
        dg[q] = 0; ell[q] = nullptr; p_lo[q] = p_hi[q] = p_self[q] = nullptr;
        if (t < T) {
          const HopLoc2 l = locate2(a.cum, a.n_hops_targets, t);
          dg[q] = a.deg[t];
          ell[q] = a.ell[l.hop] + static_cast<int64_t>(l.row) * a.k[l.hop];
        }
      }
      int s_lo[RPG], s_hi[RPG];
#pragma unroll
      for (int q = 0; q < RPG; ++q) {
        s_lo[q] = (gl < dg[q]) ? ell[q][gl] : -1;
        s_hi[q] = (gl + 8 < dg[q]) ? ell[q][gl + 8] : -1;
      }
#pragma unroll
      for (int q = 0; q < RPG; ++q) {
        const int t = tile * kTileM + group + q * (kTileM / RPG);
        if (s_lo[q] >= 0) p_lo[q] = src_row2(a, s_lo[q]);
        if (s_hi[q] >= 0) p_hi[q] = src_row2(a, s_hi[q]);
        if (t < T && gl == 0) p_self[q] = src_row2(a, t);
        p_self[q] = reinterpret_cast<const uint8_t*>(
            __shfl_sync(gmask, reinterpret_cast<unsigned long long>(p_self[q]), 0, 8));
      }
      bool waited = false;
#pragma unroll
      for (int q = 0; q < RPG; ++q) {
        const int r = group + q * (kTileM / RPG);
        const int t = tile * kTileM + r;
        float acc[NC][8];
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[c][i] = 0.f;
        for (int j0 = 0; j0 < dg[q]; j0 += 8) {
          const uint8_t* my_ptr = (j0 == 0) ? p_lo[q] : p_hi[q];
          if (j0 >= 16) {
            my_ptr = nullptr;
            if (j0 + gl < dg[q]) {
              const int s = ell[q][j0 + gl];
              if (s >= 0) my_ptr = src_row2(a, s);
            }
          }
          const int cnt = min(8, dg[q] - j0);
#pragma unroll
          for (int jb = 0; jb < 8; jb += 4) {
            uint4 v[4][NC];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const uint8_t* p = reinterpret_cast<const uint8_t*>(
                  __shfl_sync(gmask, reinterpret_cast<unsigned long long>(my_ptr), jb + jj, 8));
#pragma unroll
              for (int c = 0; c < NC; ++c)
                v[jj][c] = (jb + jj < cnt && p) ? ld_nc_v4(p + c * kChunkBytes + gl * 16) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
              for (int c = 0; c < NC; ++c) bf16x8_accum(v[jj][c], acc[c]);
          }
        }
        const float inv = dg[q] > 0 ? 1.f / static_cast<float>(dg[q]) : 0.f;
        uint4 res_mean[NC], res_self[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          res_mean[c] = pack_bf16x8(acc[c], inv);
          res_self[c] = p_self[q] ? ld_nc_v4(p_self[q] + c * kChunkBytes + gl * 16) : make_uint4(0, 0, 0, 0);
        }
        if (t < T && f.a_save) {
          uint8_t* o = reinterpret_cast<uint8_t*>(f.a_save) + static_cast<int64_t>(t) * a.d * 4;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            *reinterpret_cast<uint4*>(o + c * kChunkBytes + gl * 16) = res_mean[c];
            *reinterpret_cast<uint4*>(o + a.d * 2 + c * kChunkBytes + gl * 16) = res_self[c];
          }
        }
        // the single A buffer is free once the previous tile's MMAs have retired
        if (!waited) { mbar_wait(bar_a_empty, (it & 1) ^ 1); waited = true; }
        const uint32_t off = r * kChunkBytes + ((gl ^ (r & 7)) << 4);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          *reinterpret_cast<uint4*>(smem_a + c * kAChunkTile + off) = res_mean[c];
          *reinterpret_cast<uint4*>(smem_a + (NC + c) * kAChunkTile + off) = res_self[c];
        }
      }
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a_full);
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      const uint32_t w_bytes = NKC * N * kChunkBytes;
      mbar_expect_tx(bar_w, w_bytes);
      // one TMA bulk copy per K chunk (each <= 32 KB)
      for (int kc = 0; kc < NKC; ++kc)
        bulk_g2s(smem_u32(smem_w + kc * N * kChunkBytes),
                 reinterpret_cast<const uint8_t*>(f.w_packed) + static_cast<size_t>(kc) * N * kChunkBytes,
                 N * kChunkBytes, bar_w);
      mbar_wait(bar_w, 0);
      // instruction descriptor: D=f32, A=B=bf16, K-major both, N>>3 @17, M>>4 @24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
                             (static_cast<uint32_t>(kTileM >> 4) << 24);
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int stage = it & 1;
        const int use = it >> 1;
        mbar_wait(bar_t_empty0 + 8 * stage, (use & 1) ^ 1);
        mbar_wait(bar_a_full, it & 1);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + stage * 256;
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc) {
          const uint32_t a_base = smem_u32(smem_a + kc * kAChunkTile);
          const uint32_t b_base = smem_u32(smem_w + kc * N * kChunkBytes);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            umma_bf16(tmem_c, make_sw128_desc(a_base + k4 * 32), make_sw128_desc(b_base + k4 * 32), idesc,
                      (kc | k4) ? 1u : 0u);
          }
        }
        umma_commit(bar_a_empty);               // A buffer reusable
        umma_commit(bar_t_full0 + 8 * stage);   // accumulator ready
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue ------------------------------
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int stage = it & 1;
      const int use = it >> 1;
      mbar_wait(bar_t_full0 + 8 * stage, use & 1);
      tc_fence_after();
      const int r = quad * 32 + lane;
      const int t = tile * kTileM + r;
      __nv_bfloat16* zrow = reinterpret_cast<__nv_bfloat16*>(f.z) + static_cast<int64_t>(t) * N;
      const __nv_bfloat16* bias = reinterpret_cast<const __nv_bfloat16*>(f.bias);
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + stage * 256 + c0, v);
        if (t < T) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              x[i] = __uint_as_float(v[g * 8 + i]) + __bfloat162float(bias[c0 + g * 8 + i]);
              if (f.relu) x[i] = fmaxf(x[i], 0.f);
            }
            *reinterpret_cast<uint4*>(zrow + c0 + g * 8) = pack_bf16x8(x, 1.f);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_t_empty0 + 8 * stage);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------------------------------------
// v3: decoupled address resolution.
//
// Profiling v2 (profiles/ncu_sage_fused_v2_16producers.txt) showed the producers spending about
// half of every tile in the dependent pointer chase (deg/ELL -> node id -> owner shard) with no
// feature bytes in flight.  v3 gives the chase to four dedicated RESOLVER warps (one thread per
// tile row) that run one tile AHEAD of the loaders and publish, per row, 15 neighbour row pointers
// + the self-row pointer + the degree into a double-buffered shared-memory table.  The 16 LOADER
// warps then do nothing but stream 16-byte feature vectors (6 rows in flight per lane: the self
// row rides in slot 0 of the first batch and goes straight to the A tile) and reduce them.
// Resolver warp 0 also owns TMEM, the weight bulk copy and the tcgen05.mma issue, so the CTA
// stays at 24 warps (80 registers per thread).
//
//   warps 0-15  loaders   (8-lane group per target row, 2 rows per group per tile)
//   warps 16-19 resolvers (warp 16, lane 0: TMA weight load + MMA issue)
//   warps 20-23 epilogue  (TMEM lane quadrant = warp_id % 4)
__device__ unsigned long long g_fused_trace[148 * kFusedTraceSlots];
__device__ __forceinline__ void trace_at(unsigned long long* tr, int slot) {
  if (tr && slot < kFusedTraceSlots) tr[slot] = static_cast<unsigned long long>(clock64());
}
// timeline slots: 0 start, 1 end, 2 #tiles, then per tile i < 4 at 4 + 7 i:
//   +0 table published  +1 loaders start  +2 first row done  +3 tile loaded  +4 MMA issued
//   +5 epilogue start   +6 epilogue done
constexpr int kResolverWarps = 4;
constexpr int kThreads3 = (kProducerWarps + kResolverWarps + 4) * 32;
constexpr int kKP = 15;                      // neighbour handles published per row (+1 self)
constexpr int kTabStageBytes = kTileM * 16 * 4;   // 16 x uint32 row handles per tile row
constexpr int kStageWarpBytes = 32 * 128;    // epilogue staging: 32 rows x 64 bf16 per warp
constexpr int kBarBytes = 256;
constexpr uint32_t kNoRow = 0xFFFFFFFFu;

// A row handle is (shard << 28) | row-inside-shard; the loaders turn it into an address with one
// LDS from a 16-entry base table, so a table row costs 64 B instead of 128 B of pointers -- the
// 16 KB this frees hold the epilogue's staging buffers.
// FP8 = true: the feature table holds MXFP8 rows (d e4m3 bytes + d/32 UE8M0 block scales, padded to a multiple
// of 16 bytes; data/quantize.py) -- the loaders move HALF the bytes per row, de-quantise in registers while they
// reduce the neighbour mean in fp32 and still hand bf16 tiles to the tensor core (d = 128 only).
template <int NC, bool FP8>
__global__ void __launch_bounds__(kThreads3, 1) k_sage_fused3(SageFusedArgs f) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int N = f.n_out;
  constexpr int NKC = 2 * NC;
  constexpr int NB = 6;    // first batch: self + 5 neighbour rows in flight per lane
  constexpr int NB2 = 5;   // later batches (fan-out 10 / 15 -> 2 / 3 batches in total)
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem_w + NKC * N * kChunkBytes;
  uint8_t* tab = smem_a + NKC * kAChunkTile;                  // [2][128][16] uint32
  uint8_t* degtab = tab + 2 * kTabStageBytes;                 // [2][128] int32
  uint8_t* stg = degtab + 2 * kTileM * 4;                     // [4 warps][32 rows][128 B]
  uint8_t* bias_s = stg + 4 * kStageWarpBytes;                // [256] bf16
  uint8_t* base_s = bias_s + 512;                             // [16] shard base pointers
  uint64_t* bars = reinterpret_cast<uint64_t*>(base_s + 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t a_u32 = smem_u32(smem_a);
  const uint32_t tab_u32 = smem_u32(tab);
  const uint32_t deg_u32 = smem_u32(degtab);
  const uint32_t base_u32 = smem_u32(base_s);
  const uint32_t bar_w = smem_u32(bars + 0);
  const uint32_t bar_a_full = smem_u32(bars + 1);
  const uint32_t bar_a_empty = smem_u32(bars + 2);
  const uint32_t bar_t_full0 = smem_u32(bars + 3);    // +8: stage 1
  const uint32_t bar_t_empty0 = smem_u32(bars + 5);   // +8: stage 1
  const uint32_t bar_p_full0 = smem_u32(bars + 7);    // +8: stage 1
  const uint32_t bar_p_empty0 = smem_u32(bars + 9);   // +8: stage 1

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  unsigned long long* tr = f.trace ? f.trace + blockIdx.x * kFusedTraceSlots : nullptr;
  if (threadIdx.x == 0) trace_at(tr, 0);
  const SageAggArgs& a = f.agg;
  const int64_t row_bytes = FP8 ? static_cast<int64_t>(a.feat.row_bytes) : static_cast<int64_t>(a.d) * 2;
  // Tile order.  Low tile indices hold the early hops (largest fan-out = most neighbour rows per
  // tile); a plain round-robin hands those to the same CTAs that also get a tile of the last,
  // partial wave.  The first wave is therefore dealt in reverse: the CTAs that own an extra tile
  // start on the cheapest tiles and the expensive ones go to CTAs with one tile fewer.
  auto tile_of = [&](int it) -> int {
    return it == 0 ? static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x)
                   : it * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x);
  };
  constexpr int kResWarp0 = kProducerWarps;                   // 16
  constexpr int kEpiWarp0 = kProducerWarps + kResolverWarps;  // 20

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_a_full, kProducerWarps);
    mbar_init(bar_a_empty, 1);
    mbar_init(bar_t_full0, 1);
    mbar_init(bar_t_full0 + 8, 1);
    mbar_init(bar_t_empty0, 4);
    mbar_init(bar_t_empty0 + 8, 4);
    mbar_init(bar_p_full0, kResolverWarps);
    mbar_init(bar_p_full0 + 8, kResolverWarps);
    mbar_init(bar_p_empty0, kProducerWarps);
    mbar_init(bar_p_empty0 + 8, kProducerWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < kMaxParts) {  // shard base table (one entry for a dense local source)
    const void* b = nullptr;
    if (a.src_local)  // dense local source: "shard" p is the same matrix advanced by p * 2^28 rows, so any
                      // non-negative 32-bit row index decodes correctly
      b = reinterpret_cast<const uint8_t*>(a.src_local) +
          static_cast<int64_t>(threadIdx.x) * (int64_t(1) << 28) * row_bytes;
    else if (static_cast<int>(threadIdx.x) < a.feat.num_parts) b = a.feat.base[threadIdx.x];
    // staged remote rows (see SageFusedArgs::xcache): the slot of one (never referenced) remote part points at the cache
    if (f.xcache && static_cast<int>(threadIdx.x) == f.cache_part) b = f.xcache;
    sts64(base_u32 + threadIdx.x * 8, reinterpret_cast<uint64_t>(b));
  }
  if (warp == kResWarp0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // programmatic dependent launch: barrier init, shard base table and TMEM allocation above overlap with the tail
  // of the previous kernel; global data written by predecessors (bias, counters, ELL, rows) is only read below
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x >= 64 && threadIdx.x < 64 + (N >> 3))  // bias -> smem, 16 B per thread
    *reinterpret_cast<uint4*>(bias_s + (threadIdx.x - 64) * 16) =
        reinterpret_cast<const uint4*>(f.bias)[threadIdx.x - 64];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int T = min(a.cum[a.n_hops_targets], a.cap_targets);
  const int n_tiles = (T + kTileM - 1) / kTileM;

  if (warp < kProducerWarps) {
    // ------------------------------ loaders ------------------------------
    const int gl = lane & 7;
    const int gw = lane >> 3;
    const int group = warp * 4 + gw;  // 0..63
    constexpr int RPG = kTileM / (kProducerWarps * 4);
    auto row_addr = [&](uint32_t h) -> const uint8_t* {
      if (h == kNoRow) return nullptr;
      return reinterpret_cast<const uint8_t*>(lds64(base_u32 + ((h >> 28) << 3))) +
             static_cast<int64_t>(h & 0x0FFFFFFFu) * row_bytes;
    };
    int it = 0;
    for (int tile = tile_of(0); tile < n_tiles; tile = tile_of(++it)) {
      const int stage = it & 1;
      mbar_wait(bar_p_full0 + 8 * stage, (it >> 1) & 1);
      if (threadIdx.x == 0) trace_at(tr, 4 + 7 * it + 1);
      const uint32_t tb = tab_u32 + stage * kTabStageBytes;
      const uint32_t dgt = deg_u32 + stage * (kTileM * 4);
      bool waited = false;
      if constexpr (FP8) {
        static_assert(!FP8 || NC == 2, "the MXFP8 loader is written for d = 128");
        constexpr int NF = 8;              // rows in flight per lane in the first batch (self + 7 neighbours)
        const int cc = gl >> 2;            // bf16 K-chunk that this lane's 16 elements fall into
        const int vv = (gl & 3) * 2;       // first of the two 16-byte bf16 vectors inside that chunk
#pragma unroll 1
        for (int q = 0; q < RPG; ++q) {
          const int r = group + q * (kTileM / RPG);
          const int t = tile * kTileM + r;
          const int dg = lds32(dgt + r * 4);
          const uint32_t trow = tb + r * 64;
          const int sw = r & 15;
          const uint32_t arow = a_u32 + r * kChunkBytes;
          const uint32_t o0 = ((vv ^ (r & 7)) << 4), o1 = (((vv + 1) ^ (r & 7)) << 4);
          uint8_t* save = (t < T && f.a_save)
                              ? reinterpret_cast<uint8_t*>(f.a_save) + static_cast<int64_t>(t) * a.d * 4 + gl * 32
                              : nullptr;
          const int dgt_ = min(dg, kKP);
          float acc[16];
          {
            uint4 v[NF];
            uint32_t sc[NF];
#pragma unroll
            for (int jj = 0; jj < NF; ++jj) {
              const uint8_t* p = nullptr;
              if (jj <= dgt_) p = row_addr(static_cast<uint32_t>(lds32(trow + ((((jj + 15) & 15) ^ sw) << 2))));
              v[jj] = p ? ld_nc_v4(p + gl * 16) : make_uint4(0, 0, 0, 0);
              sc[jj] = p ? ld_nc_u32(p + a.d) : 0u;
            }
            if (!waited) { mbar_wait(bar_a_empty, (it & 1) ^ 1); waited = true; }
            {  // the row itself -> bf16 -> "self" half of the A tile
              float x[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) x[i] = 0.f;
              mxfp8x16_accum(v[0], sc[0], gl, x);
              const uint4 s0 = pack_bf16x8(x, 1.f), s1 = pack_bf16x8(x + 8, 1.f);
              sts128(arow + (NC + cc) * kAChunkTile + o0, s0);
              sts128(arow + (NC + cc) * kAChunkTile + o1, s1);
              if (save) {
                *reinterpret_cast<uint4*>(save + a.d * 2) = s0;
                *reinterpret_cast<uint4*>(save + a.d * 2 + 16) = s1;
              }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll
            for (int jj = 1; jj < NF; ++jj) mxfp8x16_accum(v[jj], sc[jj], gl, acc);
          }
          if (dgt_ > NF - 1) {             // neighbours 7..14 in one more batch
            uint4 v[NF];
            uint32_t sc[NF];
#pragma unroll
            for (int jj = 0; jj < NF; ++jj) {
              const int j = NF - 1 + jj;
              const uint8_t* p = nullptr;
              if (j < dgt_) p = row_addr(static_cast<uint32_t>(lds32(trow + ((j ^ sw) << 2))));
              v[jj] = p ? ld_nc_v4(p + gl * 16) : make_uint4(0, 0, 0, 0);
              sc[jj] = p ? ld_nc_u32(p + a.d) : 0u;
            }
#pragma unroll
            for (int jj = 0; jj < NF; ++jj) mxfp8x16_accum(v[jj], sc[jj], gl, acc);
          }
          if (dg > kKP) {  // rows wider than the table (fan-out > 15): chase the rest here
            const HopLoc2 l = locate2(a.cum, a.n_hops_targets, t);
            const int32_t* ellrow = a.ell[l.hop] + static_cast<int64_t>(l.row) * a.k[l.hop];
#pragma unroll 1
            for (int j = kKP; j < dg; ++j) {
              const int sidx = __ldg(ellrow + j);
              if (sidx < 0) continue;
              const uint8_t* p = src_row2(a, sidx);
              mxfp8x16_accum(ld_nc_v4(p + gl * 16), ld_nc_u32(p + a.d), gl, acc);
            }
          }
          const float inv = dg > 0 ? 1.f / static_cast<float>(dg) : 0.f;
          const uint4 m0 = pack_bf16x8(acc, inv), m1 = pack_bf16x8(acc + 8, inv);
          sts128(arow + cc * kAChunkTile + o0, m0);
          sts128(arow + cc * kAChunkTile + o1, m1);
          if (save) {
            *reinterpret_cast<uint4*>(save) = m0;
            *reinterpret_cast<uint4*>(save + 16) = m1;
          }
          if (threadIdx.x == 0 && q == 0) trace_at(tr, 4 + 7 * it + 2);
        }
      } else {
#pragma unroll 1
      for (int q = 0; q < RPG; ++q) {
        const int r = group + q * (kTileM / RPG);
        const int t = tile * kTileM + r;
        const int dg = lds32(dgt + r * 4);
        const uint32_t trow = tb + r * 64;
        const int sw = r & 15;
        const uint32_t off = r * kChunkBytes + ((gl ^ (r & 7)) << 4);
        uint8_t* save = (t < T && f.a_save)
                            ? reinterpret_cast<uint8_t*>(f.a_save) + static_cast<int64_t>(t) * a.d * 4 + gl * 16
                            : nullptr;
        const int dgt_ = min(dg, kKP);
        float acc[NC][8];
        {
          // first batch (peeled so the accumulators are not live across it): the row itself +
          // neighbours 0..NB-2.  Table entry 15 holds the self handle, entry j neighbour j.
          uint4 v[NB][NC];
#pragma unroll
          for (int jj = 0; jj < NB; ++jj) {
            const uint8_t* p = nullptr;
            if (jj <= dgt_) p = row_addr(static_cast<uint32_t>(lds32(trow + ((((jj + 15) & 15) ^ sw) << 2))));
#pragma unroll
            for (int c = 0; c < NC; ++c)
              v[jj][c] = p ? ld_nc_v4(p + c * kChunkBytes + gl * 16) : make_uint4(0, 0, 0, 0);
          }
          // the single A buffer is free once the previous tile's MMAs have retired; the wait
          // hides behind the loads issued above
          if (!waited) { mbar_wait(bar_a_empty, (it & 1) ^ 1); waited = true; }
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            sts128(a_u32 + (NC + c) * kAChunkTile + off, v[0][c]);
            if (save) *reinterpret_cast<uint4*>(save + a.d * 2 + c * kChunkBytes) = v[0][c];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[c][i] = 0.f;
          }
#pragma unroll
          for (int jj = 1; jj < NB; ++jj)
#pragma unroll
            for (int c = 0; c < NC; ++c) bf16x8_accum(v[jj][c], acc[c]);
        }
#pragma unroll 1
        for (int j0 = NB - 1; j0 < dgt_; j0 += NB2) {
          uint4 v[NB2][NC];
#pragma unroll
          for (int jj = 0; jj < NB2; ++jj) {
            const int j = j0 + jj;
            const uint8_t* p = nullptr;
            if (j < dgt_) p = row_addr(static_cast<uint32_t>(lds32(trow + ((j ^ sw) << 2))));
#pragma unroll
            for (int c = 0; c < NC; ++c)
              v[jj][c] = p ? ld_nc_v4(p + c * kChunkBytes + gl * 16) : make_uint4(0, 0, 0, 0);
          }
#pragma unroll
          for (int jj = 0; jj < NB2; ++jj)
#pragma unroll
            for (int c = 0; c < NC; ++c) bf16x8_accum(v[jj][c], acc[c]);
        }
        if (dg > kKP) {  // rows wider than the table (fan-out > 15): chase the rest here
          const HopLoc2 l = locate2(a.cum, a.n_hops_targets, t);
          const int32_t* ellrow = a.ell[l.hop] + static_cast<int64_t>(l.row) * a.k[l.hop];
#pragma unroll 1
          for (int j = kKP; j < dg; ++j) {
            const int sidx = __ldg(ellrow + j);
            if (sidx < 0) continue;
            const uint8_t* p = src_row2(a, sidx);
#pragma unroll
            for (int c = 0; c < NC; ++c) bf16x8_accum(ld_nc_v4(p + c * kChunkBytes + gl * 16), acc[c]);
          }
        }
        const float inv = dg > 0 ? 1.f / static_cast<float>(dg) : 0.f;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const uint4 m = pack_bf16x8(acc[c], inv);
          sts128(a_u32 + c * kAChunkTile + off, m);
          if (save) *reinterpret_cast<uint4*>(save + c * kChunkBytes) = m;
        }
        if (threadIdx.x == 0 && q == 0) trace_at(tr, 4 + 7 * it + 2);
      }
      }  // !FP8
      __syncwarp();
      if (threadIdx.x == 0) trace_at(tr, 4 + 7 * it + 3);
      if (lane == 0) mbar_arrive(bar_p_empty0 + 8 * stage);  // table stage may be refilled
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a_full);
    }
  } else if (warp < kEpiWarp0) {
    // ------------------------------ resolvers (+ MMA issue) ------------------------------
    const int r = (warp - kResWarp0) * 32 + lane;  // tile row owned by this thread
    const bool mma_thread = (warp == kResWarp0) && (lane == 0);
    if (mma_thread && tile_of(0) < n_tiles) {  // never leave a bulk copy in flight at exit
      const uint32_t w_bytes = NKC * N * kChunkBytes;
      mbar_expect_tx(bar_w, w_bytes);
      for (int kc = 0; kc < NKC; ++kc)
        bulk_g2s(smem_u32(smem_w + kc * N * kChunkBytes),
                 reinterpret_cast<const uint8_t*>(f.w_packed) + static_cast<size_t>(kc) * N * kChunkBytes,
                 N * kChunkBytes, bar_w);
    }
    // global row id (+ the node's local id in the batch) -> (shard, row in shard); with a staged remote-row cache
    // every row that is not in this GPU's HBM is read from the cache at its local id instead of from the peer
    auto handle_of = [&](int64_t gid, int local_id) -> uint32_t {
      if (gid < 0) return kNoRow;
#pragma unroll 1
      for (int p = 0; p < a.feat.num_parts; ++p)
        if (gid >= a.feat.row_begin[p] && gid < a.feat.row_begin[p + 1]) {
          if (f.xcache && !((f.local_mask >> p) & 1u))
            return (static_cast<uint32_t>(f.cache_part) << 28) | static_cast<uint32_t>(local_id);
          return (static_cast<uint32_t>(p) << 28) | static_cast<uint32_t>(gid - a.feat.row_begin[p]);
        }
      return kNoRow;
    };
    auto resolve = [&](int tile, int stage) {
      const int t = tile * kTileM + r;
      int dg = 0, kk = 0;
      const int32_t* ellp = nullptr;
      int64_t self_gid = -1;
      if (t < T) {
        const HopLoc2 l = locate2(a.cum, a.n_hops_targets, t);
        kk = min(a.k[l.hop], kKP);
        ellp = a.ell[l.hop] + static_cast<int64_t>(l.row) * a.k[l.hop];
        dg = __ldg(a.deg + t);
        if (!a.src_local) self_gid = __ldg(a.nodes + t);
      }
      int sidx[kKP];
#pragma unroll
      for (int j = 0; j < kKP; ++j) sidx[j] = (j < kk) ? __ldg(ellp + j) : -1;  // independent of dg
      const uint32_t trow = tab_u32 + stage * kTabStageBytes + r * 64;
      const int sw = r & 15;
      if (a.src_local) {
#pragma unroll
        for (int j = 0; j < kKP; ++j)
          sts32(trow + ((j ^ sw) << 2), (j < dg && sidx[j] >= 0) ? sidx[j] : static_cast<int32_t>(kNoRow));
        sts32(trow + ((15 ^ sw) << 2), (t < T) ? t : static_cast<int32_t>(kNoRow));
      } else {
        int64_t gid[kKP];
#pragma unroll
        for (int j = 0; j < kKP; ++j) gid[j] = (j < dg && sidx[j] >= 0) ? __ldg(a.nodes + sidx[j]) : -1;
        // The loaders will want these rows one tile (~7 us) from now: pull the LOCAL ones from DRAM into L2 already
        // (the loader's three dependent batches per row then cost L2 latency instead of DRAM latency).  Peer rows
        // are skipped: peer memory is not cached in the local L2.
        uint32_t hnd[kKP + 1];
#pragma unroll
        for (int j = 0; j < kKP; ++j) hnd[j] = handle_of(gid[j], sidx[j]);
        hnd[kKP] = handle_of(self_gid, t);
        if (f.l2_prefetch) {
          const unsigned pf_mask = f.local_mask | (f.xcache ? (1u << f.cache_part) : 0u);
#pragma unroll
          for (int j = 0; j <= kKP; ++j) {
            const uint32_t h = hnd[j];
            if (h != kNoRow && ((pf_mask >> (h >> 28)) & 1u)) {
              const uint8_t* p = reinterpret_cast<const uint8_t*>((f.xcache && static_cast<int>(h >> 28) == f.cache_part)
                                                                      ? f.xcache : a.feat.base[h >> 28]) +
                                 static_cast<int64_t>(h & 0x0FFFFFFFu) * row_bytes;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
              if (row_bytes > 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 128));
            }
          }
        }
#pragma unroll
        for (int j = 0; j < kKP; ++j) sts32(trow + ((j ^ sw) << 2), static_cast<int32_t>(hnd[j]));
        sts32(trow + ((15 ^ sw) << 2), static_cast<int32_t>(hnd[kKP]));
      }
      sts32(deg_u32 + (stage * kTileM + r) * 4, dg);
    };
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
                           (static_cast<uint32_t>(kTileM >> 4) << 24);
    int it = 0;
    int tile = tile_of(0);
    if (tile < n_tiles) {
      resolve(tile, 0);
      __syncwarp();
      if (mma_thread) trace_at(tr, 4 + 0);
      if (lane == 0) mbar_arrive(bar_p_full0);
    }
    for (; tile < n_tiles; tile = tile_of(++it)) {
      const int next = tile_of(it + 1);
      if (next < n_tiles) {
        const int ns = (it + 1) & 1;
        const int nu = (it + 1) >> 1;
        mbar_wait(bar_p_empty0 + 8 * ns, (nu & 1) ^ 1);
        resolve(next, ns);
        __syncwarp();
        if (mma_thread) trace_at(tr, 4 + 7 * (it + 1) + 0);
        if (lane == 0) mbar_arrive(bar_p_full0 + 8 * ns);
      }
      if (warp == kResWarp0) {
        if (lane == 0) {
          if (it == 0) mbar_wait(bar_w, 0);
          const int stage = it & 1;
          const int use = it >> 1;
          mbar_wait(bar_t_empty0 + 8 * stage, (use & 1) ^ 1);
          mbar_wait(bar_a_full, it & 1);
          tc_fence_after();
          const uint32_t tmem_c = tmem_base + stage * 256;
#pragma unroll
          for (int kc = 0; kc < NKC; ++kc) {
            const uint32_t a_base = smem_u32(smem_a + kc * kAChunkTile);
            const uint32_t b_base = smem_u32(smem_w + kc * N * kChunkBytes);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              umma_bf16(tmem_c, make_sw128_desc(a_base + k4 * 32), make_sw128_desc(b_base + k4 * 32), idesc,
                        (kc | k4) ? 1u : 0u);
            }
          }
          umma_commit(bar_a_empty);
          umma_commit(bar_t_full0 + 8 * stage);
          trace_at(tr, 4 + 7 * it + 4);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ epilogue ------------------------------
    // TMEM hands every thread one accumulator ROW (32 fp32 columns per tcgen05.ld); storing that
    // directly means 16-byte writes 512 B apart (32 cache lines and 32 half-filled sectors per
    // instruction), which the in-kernel timeline showed to be the slowest stage of the pipeline
    // (7-9 us per tile).  The tile is therefore transposed through a 4 KB per-warp staging
    // buffer: bias + ReLU + bf16 pack per row, then write-out with 8 lanes per row so that every
    // store instruction covers four full 128-byte lines.
    const int quad = warp & 3;
    const uint32_t stg_u32 = smem_u32(stg) + (warp - kEpiWarp0) * kStageWarpBytes;
    const uint32_t bias_u32 = smem_u32(bias_s);
    uint8_t* zbase = reinterpret_cast<uint8_t*>(f.z);
    int it = 0;
    for (int tile = tile_of(0); tile < n_tiles; tile = tile_of(++it)) {
      const int stage = it & 1;
      const int use = it >> 1;
      mbar_wait(bar_t_full0 + 8 * stage, use & 1);
      tc_fence_after();
      if (warp == kEpiWarp0 && lane == 0) trace_at(tr, 4 + 7 * it + 5);
      const int row0 = tile * kTileM + quad * 32;  // first tile row of this warp
      for (int c0 = 0; c0 < N; c0 += 64) {
        const int cw = min(64, N - c0);  // 64 or 32 columns in this pass
#pragma unroll 1
        for (int h = 0; h < cw; h += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + stage * 256 + c0 + h, v);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 braw;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(braw.x), "=r"(braw.y), "=r"(braw.z), "=r"(braw.w)
                         : "r"(bias_u32 + (c0 + h + g * 8) * 2));
            const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&braw);
            float x[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 bf = __bfloat1622float2(b2[i]);
              x[2 * i] = __uint_as_float(v[g * 8 + 2 * i]) + bf.x;
              x[2 * i + 1] = __uint_as_float(v[g * 8 + 2 * i + 1]) + bf.y;
            }
            if (f.relu) {
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = fmaxf(x[i], 0.f);
            }
            const int chunk = (h >> 3) + g;  // 16-byte chunk inside the staged 128-byte row
            sts128(stg_u32 + lane * 128 + ((chunk ^ (lane & 7)) << 4), pack_bf16x8(x, 1.f));
          }
        }
        if (c0 + 64 >= N) {  // all TMEM reads of this tile are done: hand the accumulator back early
          tc_fence_before();
          __syncwarp();
          if (warp == kEpiWarp0 && lane == 0) trace_at(tr, 4 + 7 * it + 6);
          if (lane == 0) mbar_arrive(bar_t_empty0 + 8 * stage);
        } else {
          __syncwarp();
        }
        // write-out: lpr lanes per row, each 16 B
        const int lpr = cw >> 3;          // 8 (64 columns) or 4 (32 columns)
        const int rpi = 32 / lpr;         // rows per store instruction
        const int rl = lane / lpr, ch = lane % lpr;
        for (int rb = 0; rb < 32; rb += rpi) {
          const int row = rb + rl;
          uint4 o;
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                       : "r"(stg_u32 + row * 128 + ((ch ^ (row & 7)) << 4)));
          const int t = row0 + row;
          if (t < T)
            *reinterpret_cast<uint4*>(zbase + (static_cast<int64_t>(t) * N + c0) * 2 + ch * 16) = o;
        }
        __syncwarp();  // staging buffer is reused by the next pass
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && tr) {
    trace_at(tr, 1);
    int cnt = 0;
    while (tile_of(cnt) < n_tiles) ++cnt;
    tr[2] = cnt;
  }
  if (warp == kResWarp0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// W [N, K] row-major bf16 -> [K/64][N][64] with the SWIZZLE_128B XOR applied, so
// the kernel can bulk-copy it verbatim into 1024-B aligned shared memory.
__global__ void k_pack_weight(const __nv_bfloat16* w, int n, int k, __nv_bfloat16* out) {
  pdl_enter();
  const int64_t total = static_cast<int64_t>(n) * k / 8;  // 16-byte vectors
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / (k / 8));
    const int v = static_cast<int>(i % (k / 8));  // 16-B vector index along K
    const int kc = v >> 3, cv = v & 7;
    const int64_t dst = (static_cast<int64_t>(kc) * n + row) * 64 + ((cv ^ (row & 7)) << 3);
    *reinterpret_cast<uint4*>(out + dst) = *reinterpret_cast<const uint4*>(w + static_cast<int64_t>(row) * k + v * 8);
  }
}

size_t fused_smem_bytes(int d, int n_out) {
  const int nkc = 2 * (d / 64);
  return static_cast<size_t>(nkc) * n_out * kChunkBytes + static_cast<size_t>(nkc) * kAChunkTile + 128 + 1024;
}

size_t fused3_smem_bytes(int d, int n_out) {
  const int nkc = 2 * (d / 64);
  return static_cast<size_t>(nkc) * n_out * kChunkBytes + static_cast<size_t>(nkc) * kAChunkTile +
         2 * kTabStageBytes + 2 * kTileM * sizeof(int32_t) + 4 * kStageWarpBytes + 512 + 128 + kBarBytes + 1024;
}

constexpr size_t kMaxSmem = 227 * 1024;

// 3 = decoupled resolver/loader kernel, 2 = the earlier monolithic-producer kernel
// (GLT_B200_FUSED_VERSION=2 selects it for A/B measurements).
int fused_version(int d, int n_out) {
  static const int forced = [] {
    const char* e = std::getenv("GLT_B200_FUSED_VERSION");
    return e ? std::atoi(e) : 0;
  }();
  if (forced == 2) return 2;
  return fused3_smem_bytes(d, n_out) <= kMaxSmem ? 3 : 2;
}

// v3 publishes 32-bit row handles (4-bit shard, 28-bit row inside the shard)
bool handles_fit(const SageAggArgs& a) {
  if (a.src_local) return true;
  if (a.feat.num_parts > 16) return false;
  for (int p = 0; p < a.feat.num_parts; ++p)
    if (a.feat.row_begin[p + 1] - a.feat.row_begin[p] >= (int64_t(1) << 28)) return false;
  return true;
}

}  // namespace

int sage_fused_supported(int d, int n_out) {
  if (d != 64 && d != 128) return 0;
  if (n_out % 32 != 0 || n_out < 32 || n_out > 256) return 0;
  return fused_smem_bytes(d, n_out) <= kMaxSmem ? 1 : 0;
}

template <int NC>
static void launch_fused_nc(const SageFusedArgs& a_in, int grid, cudaStream_t s) {
  SageFusedArgs a = a_in;
  static const bool trace_on = [] {
    const char* e = std::getenv("GLT_B200_FUSED_TRACE");
    return e && std::atoi(e) != 0;
  }();
  a.trace = nullptr;
  if (trace_on && grid <= 148) {
    // resolved per device on its first (eager) launch, i.e. outside any stream capture
    static unsigned long long* sym[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64) {
      if (sym[dev] == nullptr) {
        void* p = nullptr;
        if (cudaGetSymbolAddress(&p, g_fused_trace) == cudaSuccess) sym[dev] = reinterpret_cast<unsigned long long*>(p);
      }
      a.trace = sym[dev];
    }
  }
  if (a.feat_fp8) {
    // MXFP8 feature rows: only the decoupled (v3) kernel has the de-quantising loader, d = 128
    if constexpr (NC == 2) {
      const size_t smem = fused3_smem_bytes(a.agg.d, a.n_out);
      cudaFuncSetAttribute(k_sage_fused3<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      launch_k(k_sage_fused3<2, true>, dim3(grid), dim3(kThreads3), smem, s, a);
    }
    return;
  }
  if (fused_version(a.agg.d, a.n_out) == 3 && handles_fit(a.agg)) {
    const size_t smem = fused3_smem_bytes(a.agg.d, a.n_out);
    cudaFuncSetAttribute(k_sage_fused3<NC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    launch_k(k_sage_fused3<NC, false>, dim3(grid), dim3(kThreads3), smem, s, a);
  } else {
    const size_t smem = fused_smem_bytes(a.agg.d, a.n_out);
    cudaFuncSetAttribute(k_sage_fused<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    k_sage_fused<NC><<<grid, kThreads, smem, s>>>(a);
  }
}

void sage_fused_trace_copy(unsigned long long* host) {
  cudaMemcpyFromSymbol(host, g_fused_trace, sizeof(unsigned long long) * 148 * kFusedTraceSlots);
}

void launch_sage_fused(const SageFusedArgs& a, int num_sms, cudaStream_t s) {
  const int max_tiles = (a.agg.cap_targets + kTileM - 1) / kTileM;
  const int grid = max_tiles < num_sms ? max_tiles : num_sms;
  if (a.agg.d == 64) launch_fused_nc<1>(a, grid, s);
  else launch_fused_nc<2>(a, grid, s);
}

void launch_pack_weight(const void* w, int n, int k, void* packed, cudaStream_t s) {
  const int64_t total = static_cast<int64_t>(n) * k / 8;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  if (blocks < 1) blocks = 1;
  launch_k(k_pack_weight, dim3(blocks), dim3(256), 0, s, reinterpret_cast<const __nv_bfloat16*>(w), n, k,
           reinterpret_cast<__nv_bfloat16*>(packed));
}

}  // namespace glt
