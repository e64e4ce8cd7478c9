// Device-side helpers shared by the sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "../philox.h"
#include "glt_cuda.h"

namespace glt {

constexpr int64_t kEmptyKey = -1;

// Programmatic dependent launch (PDL).  The kernels of a training step form one long dependency chain of short
// launches; with the launch attribute set (launch_utils.h) the NEXT kernel's CTAs are scheduled as soon as every
// CTA of the current grid has called pdl_trigger(), and block in pdl_wait() until the current grid has completed
// and flushed its memory -- so launch latency, scheduling and any prologue placed before pdl_wait() overlap with
// the tail of the predecessor.  Without the attribute both instructions are no-ops.  Rule: nothing that a
// predecessor kernel may have written is read before pdl_wait().
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_trigger(); }

__device__ __forceinline__ uint32_t hash_slot(int64_t key, uint32_t mask) {
  uint64_t x = static_cast<uint64_t>(key) * 0x9E3779B97F4A7C15ULL;
  return static_cast<uint32_t>(x >> 29) & mask;
}

constexpr uint32_t kTableFull = 0xFFFFFFFFu;

// Insert `key`; returns the slot and whether this thread claimed it.  The probe sequence is bounded
// by the table size: a full table (a batch with far more unique nodes than the calibrated arena
// was sized for) yields kTableFull instead of spinning forever; callers drop that neighbour and
// count it as overflow.
__device__ __forceinline__ uint32_t table_insert(const HashTable& t, int64_t key, bool* is_new) {
  uint32_t s = hash_slot(key, t.mask);
  *is_new = false;
  for (uint32_t probes = 0; probes <= t.mask; ++probes) {
    unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(t.keys + s),
                                        static_cast<unsigned long long>(kEmptyKey),
                                        static_cast<unsigned long long>(key));
    if (prev == static_cast<unsigned long long>(kEmptyKey)) { *is_new = true; return s; }
    if (prev == static_cast<unsigned long long>(key)) return s;
    s = (s + 1) & t.mask;
  }
  return kTableFull;
}

__device__ __forceinline__ int32_t table_find_slot(const HashTable& t, int64_t key) {
  uint32_t s = hash_slot(key, t.mask);
  for (uint32_t probes = 0; probes <= t.mask; ++probes) {
    int64_t k = t.keys[s];
    if (k == key) return static_cast<int32_t>(s);
    if (k == kEmptyKey) return -1;
    s = (s + 1) & t.mask;
  }
  return -1;
}

struct RowRef {
  int64_t start;  // offset into the owning shard's indices
  int32_t deg;
  int32_t part;
};

// Owner lookup + row extent.  The shard may be peer HBM: these are plain
// ld.global on NVLink-mapped addresses, no RPC, no host staging.
__device__ __forceinline__ RowRef load_row(const GraphTable& g, int64_t v) {
  RowRef r;
  r.start = 0; r.deg = 0; r.part = -1;
#pragma unroll 1
  for (int p = 0; p < g.num_parts; ++p) {
    const int64_t b = g.parts[p].row_begin, e = g.parts[p].row_end;
    if (v >= b && v < e) {
      const int64_t* ip = g.parts[p].indptr + (v - b);
      const int64_t s0 = __ldg(ip), s1 = __ldg(ip + 1);
      r.start = s0; r.deg = static_cast<int32_t>(s1 - s0); r.part = p;
      break;
    }
  }
  return r;
}

__device__ __forceinline__ int64_t load_col(const GraphTable& g, int part, int64_t pos) {
  if (g.idx_bytes == 4) return __ldg(reinterpret_cast<const int32_t*>(g.parts[part].indices) + pos);
  return __ldg(reinterpret_cast<const int64_t*>(g.parts[part].indices) + pos);
}

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// 16-byte streaming load that bypasses L1 allocation (feature rows are touched
// once per batch; peer rows are not L2-cached locally anyway).
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ uint32_t ld_nc_u32(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// MXFP8 (OCP microscaling: e4m3 elements, one UE8M0 power-of-two scale per 32 elements).  `v` holds 16 consecutive
// elements of a row starting at element gl * 16, `scales` the row's first four block-scale bytes; the block of this
// lane's elements is gl >> 1.  acc[0..16) += dequantised values.
__device__ __forceinline__ void mxfp8x16_accum(const uint4& v, uint32_t scales, int gl, float* acc) {
  const float s = __uint_as_float(((scales >> (8 * (gl >> 1))) & 0xFFu) << 23);   // UE8M0 byte b -> 2^(b - 127)
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const __half2_raw hr = __nv_cvt_fp8x2_to_halfraw2(static_cast<__nv_fp8x2_storage_t>((w[i] >> (16 * h)) & 0xFFFFu),
                                                        __NV_E4M3);
      const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&hr));
      acc[i * 4 + h * 2] += f2.x * s;
      acc[i * 4 + h * 2 + 1] += f2.y * s;
    }
  }
}

__device__ __forceinline__ void bf16x8_accum(const uint4& v, float* acc) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    acc[2 * i] += f.x;
    acc[2 * i + 1] += f.y;
  }
}

__device__ __forceinline__ uint4 pack_bf16x8(const float* a, float scale) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(a[2 * i] * scale, a[2 * i + 1] * scale);
  return r;
}

// Row pointer of a RowTable by (already remapped) row index.
__device__ __forceinline__ const uint8_t* row_ptr(const RowTable& t, int64_t row) {
#pragma unroll 1
  for (int p = 0; p < t.num_parts; ++p) {
    if (row >= t.row_begin[p] && row < t.row_begin[p + 1])
      return reinterpret_cast<const uint8_t*>(t.base[p]) + (row - t.row_begin[p]) * t.row_bytes;
  }
  return nullptr;
}

}  // namespace glt
