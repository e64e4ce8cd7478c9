// Transposed per-batch adjacency and the atomics-free backward of the mean aggregation.
//
// EXPERIMENTAL (GraphSageEngine(use_gather_bwd=True) / GLT_B200_GATHER_BWD=1): written at the end of
// round 1 from the measured step breakdown -- zero_rows + sage_scatter_bwd (fp32 float4 atomics)
// + relu_bwd_cast cost ~78 us of the 252 us step -- and not yet the default path.
//
//   build (sampling stream, overlapped with the training of the previous batch):
//     k_tr_count   in-degree histogram of the local source ids, one histogram per hop
//     k_tr_scan*   exclusive scan over sources (3-phase, fixed grids, sizes read from the device)
//     k_tr_fill    targets appended per source, hop 0 first, so a layer that uses hops 0..h reads
//                  the first cnt_upto[h][s] entries of the segment
//   use (training stream):
//     k_sage_gather_bwd  one lane group per SOURCE row: self gradient + sum of the incoming mean
//                  gradients scaled by 1/deg(target), ReLU mask, bf16 cast, fused bias column sums.
//
// The reference has no counterpart (PyG autograd scatter on the host-built COO).
#include "device_utils.cuh"

namespace glt {

namespace {

constexpr int kScanBlock = 256;
constexpr int kScanItems = 4;  // elements per thread -> 1024 per block

__global__ void k_tr_count(TransposeArgs a, int hop) {
  const int f_begin = a.cum[hop];
  const int n_rows = min(a.cum[hop + 1] - f_begin, a.cap_rows[hop]);
  const int k = a.k[hop];
  const int64_t n = static_cast<int64_t>(n_rows) * k;
  int32_t* cnt = a.cnt + static_cast<int64_t>(hop) * a.cap_nodes;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / k), j = static_cast<int>(i % k);
    if (j >= a.deg[f_begin + r]) continue;
    const int32_t s = a.ell[hop][i];
    if (s >= 0 && s < a.cap_nodes) atomicAdd(cnt + s, 1);
  }
}

// phase 1: per-block exclusive scan of total[s] = sum_h cnt[h][s]; cnt becomes cumulative over hops
__global__ void __launch_bounds__(kScanBlock) k_tr_scan1(TransposeArgs a) {
  __shared__ int s_warp[kScanBlock / 32];
  const int base = (blockIdx.x * kScanBlock + threadIdx.x) * kScanItems;
  int v[kScanItems];
  int local = 0;
#pragma unroll
  for (int q = 0; q < kScanItems; ++q) {
    const int s = base + q;
    int tot = 0;
    if (s < a.cap_nodes) {
      for (int h = 0; h < a.n_hops; ++h) {
        int32_t* c = a.cnt + static_cast<int64_t>(h) * a.cap_nodes + s;
        tot += *c;
        *c = tot;  // cumulative over hops 0..h
      }
    }
    v[q] = local;  // exclusive inside the thread
    local += tot;
  }
  // block-wide exclusive scan of `local`
  int x = local;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_warp[w] = x;
  __syncthreads();
  if (w == 0) {
    int t = lane < kScanBlock / 32 ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    if (lane < kScanBlock / 32) s_warp[lane] = t;  // inclusive warp totals
  }
  __syncthreads();
  const int warp_off = w > 0 ? s_warp[w - 1] : 0;
  const int thread_excl = warp_off + x - local;
#pragma unroll
  for (int q = 0; q < kScanItems; ++q) {
    const int s = base + q;
    if (s < a.cap_nodes) a.off[s] = thread_excl + v[q];
  }
  if (threadIdx.x == kScanBlock - 1) a.block_sums[blockIdx.x] = warp_off + x;  // block total
}

// phase 2: one block turns the block totals into exclusive block offsets (in place) + grand total
__global__ void __launch_bounds__(1024) k_tr_scan2(TransposeArgs a, int n_blocks) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int val = i < n_blocks ? a.block_sums[i] : 0;
    int x = val;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    if (w == 0) {
      int t = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      s_warp[lane] = t;
    }
    __syncthreads();
    const int incl = (w > 0 ? s_warp[w - 1] : 0) + x;
    const int carry = s_carry;
    if (i < n_blocks) a.block_sums[i] = carry + incl - val;  // exclusive
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) a.off[a.cap_nodes] = s_carry;  // total number of transposed edges
}

// phase 3: add the block offsets
__global__ void __launch_bounds__(kScanBlock) k_tr_scan3(TransposeArgs a) {
  const int add = a.block_sums[blockIdx.x];
  const int base = (blockIdx.x * kScanBlock + threadIdx.x) * kScanItems;
#pragma unroll
  for (int q = 0; q < kScanItems; ++q)
    if (base + q < a.cap_nodes) a.off[base + q] += add;
}

__global__ void k_tr_fill(TransposeArgs a, int hop) {
  const int f_begin = a.cum[hop];
  const int n_rows = min(a.cum[hop + 1] - f_begin, a.cap_rows[hop]);
  const int k = a.k[hop];
  const int64_t n = static_cast<int64_t>(n_rows) * k;
  // entries of hop h start after the entries of the earlier hops: offset by the cumulative count
  const int32_t* before = hop > 0 ? a.cnt + static_cast<int64_t>(hop - 1) * a.cap_nodes : nullptr;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / k), j = static_cast<int>(i % k);
    if (j >= a.deg[f_begin + r]) continue;
    const int32_t s = a.ell[hop][i];
    if (s < 0 || s >= a.cap_nodes) continue;
    const int pos = a.off[s] + (before ? before[s] : 0) + atomicAdd(a.cursor + static_cast<int64_t>(hop) * a.cap_nodes + s, 1);
    if (pos < a.cap_edges) a.tgt[pos] = f_begin + r;
  }
}

inline int grid_for(int64_t items, int per_block, int max_blocks = 148 * 16) {
  int64_t b = (items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return static_cast<int>(b);
}

// ------------------------------------------------------------------------------------------
template <int LPR, int VPL>
__global__ void __launch_bounds__(256) k_sage_gather_bwd(SageGatherBwdArgs a) {
  extern __shared__ float s_col[];  // [d] block partial of the bias gradient
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR;
  const int gw = lane / LPR;
  const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (gw * LPR));
  const int T = min(a.cum[a.n_hops_targets], a.cap_targets);
  const int S = min(a.cum[a.n_hops_targets + 1], a.cap_src);
  const int wpb = blockDim.x >> 5;
  const int nvec = a.d >> 3;
  if (a.colsum)
    for (int c = threadIdx.x; c < a.d; c += blockDim.x) s_col[c] = 0.f;
  __syncthreads();
  // per-thread running column sums of the bias gradient (a shared-memory-atomic-per-row variant was measured at
  // 2.4x the kernel time: every row of a CTA hits the same d addresses)
  float csum[VPL][8];
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) csum[v][i] = 0.f;
  const uint8_t* dA = reinterpret_cast<const uint8_t*>(a.dA);
  const int64_t a_row = static_cast<int64_t>(a.d) * 4;  // bytes of one [mean | self] row
  for (int base = (blockIdx.x * wpb + (threadIdx.x >> 5)) * RPW; base < a.cap_src; base += gridDim.x * wpb * RPW) {
    const int s = base + gw;
    if (s >= a.cap_src) continue;
    uint8_t* o = reinterpret_cast<uint8_t*>(a.dPre) + static_cast<int64_t>(s) * a.d * 2;
    if (s >= S) {  // rows beyond the batch: the GEMMs run over the arena capacity
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int c = v * LPR + gl;
        if (c < nvec) reinterpret_cast<uint4*>(o)[c] = make_uint4(0, 0, 0, 0);
      }
      continue;
    }
    float acc[VPL][8];
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[v][i] = 0.f;
    if (s < T) {  // gradient through the "self" half of A
      const uint8_t* g = dA + static_cast<int64_t>(s) * a_row + a.d * 2;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int c = v * LPR + gl;
        if (c < nvec) bf16x8_accum(ld_nc_v4(g + c * 16), acc[v]);
      }
    }
    const int n_in = a.cnt_upto[s];
    const int e0 = a.off[s];
    for (int j0 = 0; j0 < n_in; j0 += LPR) {
      int my_t = -1;
      float my_inv = 0.f;
      if (j0 + gl < n_in) {
        my_t = a.tgt[e0 + j0 + gl];
        const int dg = a.deg[my_t];
        my_inv = dg > 0 ? 1.f / static_cast<float>(dg) : 0.f;
      }
      const int cnt = min(LPR, n_in - j0);
      for (int jj = 0; jj < cnt; ++jj) {
        const int t = __shfl_sync(gmask, my_t, jj, LPR);
        const float inv = __shfl_sync(gmask, my_inv, jj, LPR);
        if (t < 0 || t >= T) continue;
        const uint8_t* g = dA + static_cast<int64_t>(t) * a_row;  // "mean" half
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int c = v * LPR + gl;
          if (c < nvec) {
            float tmp[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) tmp[i] = 0.f;
            bf16x8_accum(ld_nc_v4(g + c * 16), tmp);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[v][i] += tmp[i] * inv;
          }
        }
      }
    }
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = v * LPR + gl;
      if (c >= nvec) continue;
      if (a.Z) {
        float z[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = 0.f;
        bf16x8_accum(ld_nc_v4(reinterpret_cast<const uint8_t*>(a.Z) + static_cast<int64_t>(s) * a.d * 2 + c * 16), z);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (!(z[i] > 0.f)) acc[v][i] = 0.f;
      }
      const uint4 packed = pack_bf16x8(acc[v], 1.f);
      reinterpret_cast<uint4*>(o)[c] = packed;
      if (a.colsum) {  // sum exactly what the GEMMs will see (bf16-rounded)
        float r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = 0.f;
        bf16x8_accum(packed, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) csum[v][i] += r[i];
      }
    }
  }
  if (a.colsum) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = v * LPR + gl;
      if (c < nvec) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (csum[v][i] != 0.f) atomicAdd(s_col + c * 8 + i, csum[v][i]);
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < a.d; c += blockDim.x)
      if (s_col[c] != 0.f) atomicAdd(a.colsum + c, s_col[c]);
  }
}

}  // namespace

void launch_build_transpose(const TransposeArgs& a, cudaStream_t s) {
  cudaMemsetAsync(a.cnt, 0, sizeof(int32_t) * static_cast<size_t>(a.n_hops) * a.cap_nodes, s);
  cudaMemsetAsync(a.cursor, 0, sizeof(int32_t) * static_cast<size_t>(a.n_hops) * a.cap_nodes, s);
  for (int h = 0; h < a.n_hops; ++h)
    k_tr_count<<<grid_for(static_cast<int64_t>(a.cap_rows[h]) * a.k[h], 256 * 4), 256, 0, s>>>(a, h);
  const int per_block = kScanBlock * kScanItems;
  const int n_blocks = (a.cap_nodes + per_block - 1) / per_block;
  k_tr_scan1<<<n_blocks, kScanBlock, 0, s>>>(a);
  k_tr_scan2<<<1, 1024, 0, s>>>(a, n_blocks);
  k_tr_scan3<<<n_blocks, kScanBlock, 0, s>>>(a);
  for (int h = 0; h < a.n_hops; ++h)
    k_tr_fill<<<grid_for(static_cast<int64_t>(a.cap_rows[h]) * a.k[h], 256 * 4), 256, 0, s>>>(a, h);
}

// Lane-group width of the gather backward.  Measured on B200 (products shape, 54 k source rows x 256, 66 k in-edges):
// 32 lanes per row = 65 us for the layer-2 launch; 8 lanes x 4 vectors (4 rows per warp in flight) was slower still
// (pipelined step 0.435 ms vs 0.268 ms), as was accumulating the bias column sums with shared-memory atomics.  The
// fp32-atomic scatter path (zero_rows + sage_scatter_bwd + relu_bwd_cast = 39 us for the same layer) therefore stays
// the engine default; this kernel is kept, tested, behind use_gather_bwd / GLT_B200_GATHER_BWD=1.
#define GLT_DISPATCH_WIDTH_T(D, ...)                                          \
  do {                                                                        \
    const int nvec_ = (D) / 8;                                                \
    if (nvec_ <= 4) { constexpr int LPR = 4, VPL = 1; __VA_ARGS__; }          \
    else if (nvec_ <= 8) { constexpr int LPR = 8, VPL = 1; __VA_ARGS__; }     \
    else if (nvec_ <= 16) { constexpr int LPR = 16, VPL = 1; __VA_ARGS__; }   \
    else if (nvec_ <= 32) { constexpr int LPR = 32, VPL = 1; __VA_ARGS__; }   \
    else if (nvec_ <= 64) { constexpr int LPR = 32, VPL = 2; __VA_ARGS__; }   \
    else { constexpr int LPR = 32, VPL = 4; __VA_ARGS__; }                    \
  } while (0)

void launch_sage_gather_bwd(const SageGatherBwdArgs& a, cudaStream_t s) {
  if (a.colsum) cudaMemsetAsync(a.colsum, 0, sizeof(float) * a.d, s);
  GLT_DISPATCH_WIDTH_T(a.d, {
    k_sage_gather_bwd<LPR, VPL><<<grid_for(a.cap_src, 8 * (32 / LPR), 148 * 8), 256,
                                 a.colsum ? sizeof(float) * a.d : 0, s>>>(a);
  });
}

}  // namespace glt
