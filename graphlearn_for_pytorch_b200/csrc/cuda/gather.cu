// Multi-source row gather (UnifiedTensor / Feature lookup).
//
// out[i,:] = part(owner(idx[i]))[idx[i] - begin(owner)], where a part is local
// HBM, a peer GPU's HBM mapped over NVLink (CUDA IPC) or pinned host memory.
// Capability parity with the reference's GatherTensorKernel
// (csrc/cuda/unified_tensor.cu:47-81: warp per row, scalar T loads); here rows
// move as 16-byte vectors with L1::no_allocate, several rows share a warp when
// rows are narrow, every lane keeps up to four loads in flight, and the row
// count can stay on the device (n_dev) so the launch is CUDA-graph capturable.
#include "device_utils.cuh"
#include "launch_utils.h"

namespace glt {

namespace {

// LPR lanes cooperate on one row; every group keeps R independent rows in flight: the id ->
// (id2index) -> owner-shard pointer chase and the 16-byte row loads of R rows are issued back to
// back before anything is stored.  Rows in peer HBM cost ~2 us per dependent step over NVLink, so
// the kernel is a latency x parallelism product: R = 4 quadruples the bytes in flight per SM for
// the common 256/512-byte rows (measured effect on peer-HBM gather in profiles/).
template <int LPR, int R>
__global__ void __launch_bounds__(256) k_gather_vec(RowTable t, const int64_t* idx,
                                                    const int64_t* id2index, int64_t n,
                                                    const int32_t* n_dev, uint8_t* out,
                                                    int64_t out_row_bytes, int64_t map_len) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR;
  const int gw = lane / LPR;
  const int64_t n_valid = n_dev ? min(static_cast<int64_t>(*n_dev), n) : n;
  const int nvec = static_cast<int>(t.row_bytes >> 4);
  const int64_t wpb = blockDim.x >> 5;
  for (int64_t base = (blockIdx.x * wpb + (threadIdx.x >> 5)) * (RPW * R); base < n;
       base += static_cast<int64_t>(gridDim.x) * wpb * (RPW * R)) {
    int64_t row[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int64_t r = base + k * RPW + gw;
      row[k] = (r < n_valid) ? idx[r] : -1;
    }
    if (id2index) {
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (row[k] >= 0) row[k] = (map_len <= 0 || row[k] < map_len) ? id2index[row[k]] : -1;
    }
    const uint4* src[R];
#pragma unroll
    for (int k = 0; k < R; ++k)
      src[k] = row[k] >= 0 ? reinterpret_cast<const uint4*>(row_ptr(t, row[k])) : nullptr;
    for (int c = gl; c < nvec; c += LPR) {
      uint4 v[R];
#pragma unroll
      for (int k = 0; k < R; ++k) v[k] = src[k] ? ld_nc_v4(src[k] + c) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int64_t r = base + k * RPW + gw;
        if (r < n) reinterpret_cast<uint4*>(out + r * out_row_bytes)[c] = v[k];
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_gather_scalar(RowTable t, const int64_t* idx,
                                                       const int64_t* id2index, int64_t n,
                                                       const int32_t* n_dev, uint8_t* out,
                                                       int64_t out_row_bytes, int64_t map_len) {
  const int lane = threadIdx.x & 31;
  const int64_t n_valid = n_dev ? min(static_cast<int64_t>(*n_dev), n) : n;
  const int nel = static_cast<int>(t.row_bytes / sizeof(T));
  const int64_t wpb = blockDim.x >> 5;
  for (int64_t r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n;
       r += static_cast<int64_t>(gridDim.x) * wpb) {
    const uint8_t* src = nullptr;
    if (r < n_valid) {
      int64_t row = idx[r];
      if (row >= 0 && id2index) row = (map_len <= 0 || row < map_len) ? id2index[row] : -1;
      if (row >= 0) src = row_ptr(t, row);
    }
    T* dst = reinterpret_cast<T*>(out + r * out_row_bytes);
    const T* s = reinterpret_cast<const T*>(src);
    for (int c = lane; c < nel; c += 32) dst[c] = src ? s[c] : T(0);
  }
}

__global__ void k_gather_i64(RowTable t, const int64_t* idx, int64_t n, const int32_t* n_dev,
                             int64_t* out) {
  const int64_t n_valid = n_dev ? min(static_cast<int64_t>(*n_dev), n) : n;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t v = -1;
    if (i < n_valid && idx[i] >= 0) {
      const uint8_t* p = row_ptr(t, idx[i]);
      if (p) v = *reinterpret_cast<const int64_t*>(p);
    }
    out[i] = v;
  }
}

inline int grid_for(int64_t items, int per_block, int max_blocks = 148 * 16) {
  int64_t b = (items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return static_cast<int>(b);
}

// LPR lanes per row, R rows in flight per lane group (same latency x parallelism reasoning as k_gather_vec): only rows
// that live outside this GPU's HBM are copied.
template <int LPR, int R>
__global__ void __launch_bounds__(256) k_stage_remote_rows(RowTable t, unsigned local_mask, const int64_t* nodes,
                                                           const int32_t* cum, int n_idx, int cap_nodes, uint8_t* xcache) {
  pdl_enter();
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR, gw = lane / LPR;
  const int n = min(cum[n_idx], cap_nodes);
  const int nvec = static_cast<int>(t.row_bytes >> 4);
  const int wpb = blockDim.x >> 5;
  for (int base = (blockIdx.x * wpb + (threadIdx.x >> 5)) * (RPW * R); base < n; base += gridDim.x * wpb * (RPW * R)) {
    const uint4* src[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int s = base + k * RPW + gw;
      src[k] = nullptr;
      if (s < n) {
        const int64_t gid = nodes[s];
#pragma unroll 1
        for (int p = 0; p < t.num_parts; ++p)
          if (gid >= t.row_begin[p] && gid < t.row_begin[p + 1]) {
            if (!((local_mask >> p) & 1u))
              src[k] = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(t.base[p]) +
                                                      (gid - t.row_begin[p]) * t.row_bytes);
            break;
          }
      }
    }
    for (int c = gl; c < nvec; c += LPR) {
      uint4 v[R];
#pragma unroll
      for (int k = 0; k < R; ++k) v[k] = src[k] ? ld_nc_v4(src[k] + c) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (src[k]) reinterpret_cast<uint4*>(xcache + static_cast<int64_t>(base + k * RPW + gw) * t.row_bytes)[c] = v[k];
    }
  }
}

// 8 lanes per row, 16 e4m3 elements per lane and iteration (d multiple of 128 keeps every lane busy)
__global__ void __launch_bounds__(256) k_gather_mxfp8(RowTable t, const int64_t* idx, int64_t n, int d,
                                                      __nv_bfloat16* out) {
  const int lane = threadIdx.x & 31;
  const int gl = lane & 7, gw = lane >> 3;
  const int64_t wpb = blockDim.x >> 5;
  for (int64_t base = (blockIdx.x * wpb + (threadIdx.x >> 5)) * 4; base < n;
       base += static_cast<int64_t>(gridDim.x) * wpb * 4) {
    const int64_t r = base + gw;
    if (r >= n) continue;
    const int64_t row = idx[r];
    const uint8_t* p = row >= 0 ? row_ptr(t, row) : nullptr;
    for (int e0 = gl * 16; e0 < d; e0 += 128) {
      float x[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = 0.f;
      if (p) {
        // scale bytes of this 128-element span: 4 consecutive bytes at d + e0 / 32 (e0 / 32 is a multiple of 4
        // for the span start; the lane picks its block inside mxfp8x16_accum from gl)
        const uint32_t sc = ld_nc_u32(p + d + ((e0 - gl * 16) >> 5));
        mxfp8x16_accum(ld_nc_v4(p + e0), sc, gl, x);
      }
      __nv_bfloat16* o = out + r * d + e0;
      *reinterpret_cast<uint4*>(o) = pack_bf16x8(x, 1.f);
      *reinterpret_cast<uint4*>(o + 8) = pack_bf16x8(x + 8, 1.f);
    }
  }
}

}  // namespace

void launch_gather_rows(RowTable t, const int64_t* idx, const int64_t* id2index, int64_t n,
                        const int32_t* n_dev, void* out, int64_t out_row_bytes, cudaStream_t s,
                        int64_t id2index_len) {
  if (n <= 0) return;
  uint8_t* o = reinterpret_cast<uint8_t*>(out);
  bool aligned = (t.row_bytes % 16 == 0) && (out_row_bytes % 16 == 0) &&
                 (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  for (int p = 0; p < t.num_parts; ++p) aligned &= (reinterpret_cast<uintptr_t>(t.base[p]) % 16 == 0);
  if (aligned) {
    const int nvec = static_cast<int>(t.row_bytes / 16);
  // small lookups spread one row per group over as many SMs as possible; large ones keep four rows
  // in flight per group
#define GLT_LAUNCH_VEC(LPR)                                                                            \
  do {                                                                                                 \
    if (n >= static_cast<int64_t>(148) * 2 * 8 * (32 / LPR) * 4)                                       \
      k_gather_vec<LPR, 4><<<grid_for(n, 8 * (32 / LPR) * 4), 256, 0, s>>>(t, idx, id2index, n, n_dev, \
                                                                            o, out_row_bytes,          \
                                                                            id2index_len);             \
    else                                                                                               \
      k_gather_vec<LPR, 1><<<grid_for(n, 8 * (32 / LPR)), 256, 0, s>>>(t, idx, id2index, n, n_dev, o,  \
                                                                        out_row_bytes, id2index_len);  \
  } while (0)
    if (nvec <= 1) GLT_LAUNCH_VEC(1);
    else if (nvec <= 2) GLT_LAUNCH_VEC(2);
    else if (nvec <= 4) GLT_LAUNCH_VEC(4);
    else if (nvec <= 8) GLT_LAUNCH_VEC(8);
    else if (nvec <= 16) GLT_LAUNCH_VEC(16);
    else GLT_LAUNCH_VEC(32);
#undef GLT_LAUNCH_VEC
    return;
  }
  const int g = grid_for(n, 8);
  if (t.row_bytes % 4 == 0 && out_row_bytes % 4 == 0)
    k_gather_scalar<uint32_t><<<g, 256, 0, s>>>(t, idx, id2index, n, n_dev, o, out_row_bytes, id2index_len);
  else if (t.row_bytes % 2 == 0 && out_row_bytes % 2 == 0)
    k_gather_scalar<uint16_t><<<g, 256, 0, s>>>(t, idx, id2index, n, n_dev, o, out_row_bytes, id2index_len);
  else
    k_gather_scalar<uint8_t><<<g, 256, 0, s>>>(t, idx, id2index, n, n_dev, o, out_row_bytes, id2index_len);
}

// NVSwitch multicast store: one write lands in the replica of every GPU bound to the
// multicast object (hot-cache fill).  `mc_dst` is a multicast virtual address.
__global__ void k_multimem_copy(const uint4* __restrict__ src, uint4* mc_dst, int64_t n16) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n16;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint4 v = src[i];
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + i),
                 "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)),
                 "f"(__uint_as_float(v.w))
                 : "memory");
  }
}

void launch_multimem_copy(const void* src, void* mc_dst, int64_t nbytes, cudaStream_t s) {
  const int64_t n16 = nbytes / 16;
  if (n16 <= 0) return;
  int64_t blocks = (n16 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_multimem_copy<<<static_cast<int>(blocks), 256, 0, s>>>(reinterpret_cast<const uint4*>(src),
                                                           reinterpret_cast<uint4*>(mc_dst), n16);
}

void launch_gather_i64(RowTable t, const int64_t* idx, int64_t n, const int32_t* n_dev,
                       int64_t* out, cudaStream_t s) {
  if (n <= 0) return;
  k_gather_i64<<<grid_for(n, 256), 256, 0, s>>>(t, idx, n, n_dev, out);
}

void launch_gather_mxfp8(RowTable t, const int64_t* idx, int64_t n, int d, void* out, cudaStream_t s) {
  if (n <= 0) return;
  int64_t blocks = (n + 31) / 32;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_gather_mxfp8<<<static_cast<int>(blocks), 256, 0, s>>>(t, idx, n, d, reinterpret_cast<__nv_bfloat16*>(out));
}

void launch_stage_remote_rows(RowTable t, unsigned local_mask, const int64_t* nodes, const int32_t* cum, int n_idx,
                              int cap_nodes, void* xcache, cudaStream_t s) {
  if (cap_nodes <= 0) return;
  int blocks = (cap_nodes + 127) / 128;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (t.row_bytes <= 128)
    launch_k(k_stage_remote_rows<8, 4>, dim3(blocks), dim3(256), 0, s, t, local_mask, nodes, cum, n_idx, cap_nodes,
             reinterpret_cast<uint8_t*>(xcache));
  else
    launch_k(k_stage_remote_rows<16, 4>, dim3(blocks), dim3(256), 0, s, t, local_mask, nodes, cum, n_idx, cap_nodes,
             reinterpret_cast<uint8_t*>(xcache));
}

}  // namespace glt
