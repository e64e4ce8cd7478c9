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
#include "device_utils.cuh"

namespace glt {

namespace {

constexpr int kTileM = 128;
constexpr int kProducerWarps = 16;
constexpr int kMmaWarp = 16;
constexpr int kThreads = 21 * 32;
constexpr int kChunkBytes = 128;                    // 64 bf16 = one SWIZZLE_128B atom row
constexpr int kAChunkTile = kTileM * kChunkBytes;   // 16 KB per K-chunk of an A tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  // K-major, SWIZZLE_128B: start>>4 | LBO=1 (ignored) | SBO=1024 B | version=1 | layout=2
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

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

// W [N, K] row-major bf16 -> [K/64][N][64] with the SWIZZLE_128B XOR applied, so
// the kernel can bulk-copy it verbatim into 1024-B aligned shared memory.
__global__ void k_pack_weight(const __nv_bfloat16* w, int n, int k, __nv_bfloat16* out) {
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

}  // namespace

int sage_fused_supported(int d, int n_out) {
  if (d != 64 && d != 128) return 0;
  if (n_out % 32 != 0 || n_out < 32 || n_out > 256) return 0;
  return fused_smem_bytes(d, n_out) <= 227 * 1024 ? 1 : 0;
}

void launch_sage_fused(const SageFusedArgs& a, int num_sms, cudaStream_t s) {
  const int d = a.agg.d;
  const size_t smem = fused_smem_bytes(d, a.n_out);
  const int max_tiles = (a.agg.cap_targets + kTileM - 1) / kTileM;
  const int grid = max_tiles < num_sms ? max_tiles : num_sms;
  if (d == 64) {
    cudaFuncSetAttribute(k_sage_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    k_sage_fused<1><<<grid, kThreads, smem, s>>>(a);
  } else {
    cudaFuncSetAttribute(k_sage_fused<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    k_sage_fused<2><<<grid, kThreads, smem, s>>>(a);
  }
}

void launch_pack_weight(const void* w, int n, int k, void* packed, cudaStream_t s) {
  const int64_t total = static_cast<int64_t>(n) * k / 8;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  if (blocks < 1) blocks = 1;
  k_pack_weight<<<blocks, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(w), n, k,
                                       reinterpret_cast<__nv_bfloat16*>(packed));
}

}  // namespace glt
