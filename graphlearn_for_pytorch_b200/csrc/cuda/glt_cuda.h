// Host-visible declarations of the sm_100a kernels (torch-free: raw pointers and
// a cudaStream_t) so the .cu files compile in seconds and the binding layer is
// the only translation unit that sees torch headers.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace glt {

constexpr int kMaxParts = 16;  // 8 GPUs of one NVSwitch box + optional host tiers

// One CSR shard of a range-partitioned graph.  Pointers may live in local HBM,
// in a peer GPU's HBM (mapped over NVLink via CUDA IPC) or in pinned host
// memory; the kernels dereference them directly -- this is the generalisation
// of the reference's pointer-table gather (csrc/cuda/unified_tensor.cu:35-81)
// to graph topology.
struct CsrShard {
  const int64_t* indptr;   // [row_end - row_begin + 1], shard-local offsets
  const void* indices;     // int32 or int64 global column ids
  const int64_t* eids;     // optional global edge ids
  const float* weights;    // optional edge weights
  int64_t row_begin;       // first global row id owned by this shard
  int64_t row_end;         // one past the last owned row
};

struct GraphTable {
  int num_parts;
  int idx_bytes;  // 4: int32 column ids, 8: int64
  CsrShard parts[kMaxParts];
};

// Range-partitioned row store (features / labels).  Part p holds rows
// [row_begin[p], row_begin[p+1]).
struct RowTable {
  int num_parts;
  int64_t row_bytes;
  const void* base[kMaxParts];
  int64_t row_begin[kMaxParts + 1];
};

// Device-resident id -> local-id hash table (open addressing, linear probe).
struct HashTable {
  int64_t* keys;   // capacity slots, empty = -1
  int32_t* vals;   // dense local id of the key in this slot
  int32_t* aux;    // first-occurrence index (ordered insert only)
  uint32_t mask;   // capacity - 1
};

// Per-batch device-side bookkeeping; the host never reads it on the hot path.
//   cum[h]      : number of local nodes after hop h-1 (cum[0] = 0, cum[1] = #seeds)
//   edges[h]    : number of sampled edges of hop h
struct BatchCounters {
  int32_t* cum;     // [max_hops + 2]
  int32_t* edges;   // [max_hops]
  int32_t* cursor;  // running number of local nodes
  int32_t* overflow;  // optional: neighbours dropped by the capacity guard (cumulative)
};

// ---- sampling.cu -----------------------------------------------------------
void launch_table_clear(HashTable t, cudaStream_t s);

// Ordered (first-occurrence) insertion of the seed ids.  Writes nodes[0..n0),
// seed_local[i] (local id of seeds[i]), cum[0]=0, cum[1]=n0, cursor=n0.
void launch_init_seeds(const int64_t* seeds, int n_seeds, const int32_t* n_seeds_dev,
                       HashTable t, int64_t* nodes, int32_t* seed_local, int32_t* scratch,
                       BatchCounters c, int32_t* step_dev, int step_inc, cudaStream_t s);

struct HopArgs {
  GraphTable g;
  HashTable t;
  BatchCounters c;
  int64_t* nodes;       // local -> global id, appended by this hop
  int32_t* ell;         // [cap_rows * k] slot ids -> local ids after relabel; -1 = empty
  int64_t* ell_eids;    // optional [cap_rows * k]
  int32_t* deg;         // [cap_nodes] sampled degree per target local id
  int hop;              // reads frontier [cum[hop], cum[hop+1])
  int k;                // fanout (>0)
  int cap_rows;         // worst-case frontier size
  int cap_nodes;        // node arena capacity (overflow guard)
  int cap_rows_next;    // capacity of the next hop's frontier
  int weighted;         // 1: exponential-race weighted sampling
  int replace;          // 1: with replacement
  uint64_t seed;
  uint32_t stream;
  const int32_t* stream_dev;  // optional device step counter mixed into the Philox stream
  // --- heterogeneous graphs: frontier and neighbours belong to different node types -----------------
  int64_t* nodes_out;         // id list of the NEIGHBOUR type (appended); nullptr = same list as `nodes`
  const int32_t* bound_ptr;   // precomputed id bound of the neighbour type for this hop (k_hetero_finalize);
                              // nullptr = derive it from c.cum / c.cursor (single node type)
  uint32_t stream_stride;     // Philox stream ids consumed per device step (0 = 8, the homogeneous arena)
};
void launch_sample_hop(const HopArgs& a, cudaStream_t s);
void launch_relabel_hop(const HopArgs& a, cudaStream_t s);
// deterministic local-id order of a hop's new nodes (see sampling.cu): keys -> (caller sorts tmp) -> assign
void launch_det_keys(const HopArgs& a, int64_t* tmp, cudaStream_t s);
void launch_det_assign(const HopArgs& a, const int64_t* sorted, cudaStream_t s);

// Heterogeneous sampling (native hetero inducer; reference csrc/cuda/inducer.cu:194-338 CUDAHeteroInducer +
// python/sampler/neighbor_sampler.py:232-317): one GROUPED launch per hop over all relations.  `descs` is a
// device array of `n_rel` HopArgs (blockIdx.y selects the relation; relations with k <= 0 are skipped);
// `max_k` / `max_rows` are the maxima over the relations (template / grid selection).
void launch_sample_hop_grouped(const HopArgs* descs, int n_rel, int max_k, int max_rows, cudaStream_t s);
void launch_relabel_hop_grouped(const HopArgs* descs, int n_rel, int max_k, int max_rows, cudaStream_t s);
struct HeteroTypeState {
  int32_t* cum;       // [6] of this node type
  int32_t* cursor;
  int cap_nodes;
  int cap_rows[5];    // frontier capacity per hop (+ trailing entry)
};
// cum[hop + 2] = cursor = min(cursor, cap_nodes, cum[hop + 1] + cap_rows[hop + 1]) for every node type
void launch_hetero_finalize(const HeteroTypeState* types, int n_types, int hop, cudaStream_t s);

// One-hop API sampler (NeighborOutput): fixed-stride output, no dedup.
void launch_sample_one_hop(GraphTable g, const int64_t* seeds, int n, int k, int weighted,
                           int replace, uint64_t seed, uint32_t stream, int64_t* out_nbrs,
                           int64_t* out_eids, int32_t* out_cnt, cudaStream_t s);
// Degrees of arbitrary rows (all-neighbour fanout sizing).
void launch_lookup_degree(GraphTable g, const int64_t* ids, int n, int64_t* out, cudaStream_t s);
// Full-neighbourhood copy: out[offs[i] .. offs[i+1]) = N(ids[i]).
void launch_copy_neighbors(GraphTable g, const int64_t* ids, int n, const int64_t* offs,
                           int64_t* out_nbrs, int64_t* out_eids, cudaStream_t s);

// ELL -> COO of one hop with exact offsets (exclusive scan of deg over the frontier).
void launch_ell_to_coo(const int32_t* ell, const int64_t* ell_eids, const int32_t* deg,
                       const int64_t* offs, const int32_t* cum, int hop, int k, int cap_rows,
                       int64_t* rows, int64_t* cols, int64_t* eids, cudaStream_t s);

// Generic table ops for the inducer API (unordered insert + lookup).
void launch_table_insert(HashTable t, const int64_t* keys, int64_t n, int64_t* nodes,
                         int32_t* cursor, int cap_nodes, int32_t* out_slots, cudaStream_t s);
void launch_table_resolve(HashTable t, int32_t* slots_inout, int64_t n, cudaStream_t s);
void launch_table_lookup(HashTable t, const int64_t* keys, int64_t n, int32_t* out, cudaStream_t s);

// ---- negative / subgraph / walk / prob (graph_ops.cu) ------------------------
void launch_negative_sample(GraphTable g, int64_t num_rows, int64_t num_cols, int req, int trials,
                            int padding, uint64_t seed, uint32_t stream, int64_t* out_rows,
                            int64_t* out_cols, int32_t* out_count, cudaStream_t s);
void launch_subgraph_count(GraphTable g, HashTable t, const int64_t* nodes, int n, int64_t* cnt,
                           cudaStream_t s);
void launch_subgraph_fill(GraphTable g, HashTable t, const int64_t* nodes, int n,
                          const int64_t* offs, int64_t* rows, int64_t* cols, int64_t* eids,
                          cudaStream_t s);
void launch_random_walk(GraphTable g, const int64_t* starts, int n, int walk_len, float p, float q,
                        uint64_t seed, uint32_t stream, int64_t* out, cudaStream_t s);
void launch_nbr_prob(GraphTable g, GraphTable nbr_g, const float* last, const float* nbr_last,
                     int64_t n, int64_t n_nbr, int k, float* cur, cudaStream_t s);

// ---- gather.cu ---------------------------------------------------------------
// out[i, :] = table[idx[i]] (idx optionally remapped through id2index); rows
// with idx < 0 or i >= *n_dev (when n_dev != nullptr) are zero-filled.
// id2index_len > 0: ids outside [0, id2index_len) produce zero rows instead of an out-of-bounds read
void launch_gather_rows(RowTable t, const int64_t* idx, const int64_t* id2index, int64_t n,
                        const int32_t* n_dev, void* out, int64_t out_row_bytes, cudaStream_t s,
                        int64_t id2index_len = 0);
void launch_gather_i64(RowTable t, const int64_t* idx, int64_t n, const int32_t* n_dev,
                       int64_t* out, cudaStream_t s);
// copy `nbytes` (multiple of 16) from local memory to an NVSwitch multicast address
void launch_multimem_copy(const void* src, void* mc_dst, int64_t nbytes, cudaStream_t s);
// out[i, 0..d) (bf16) = dequantised MXFP8 row idx[i] of `t` (rows of d + 16 bytes, see data/quantize.py)
void launch_gather_mxfp8(RowTable t, const int64_t* idx, int64_t n, int d, void* out, cudaStream_t s);
// xcache[s, :] = row nodes[s] of `t` for every local id s < cum[n_idx] whose row is NOT in a part of `local_mask`
// (peer HBM / host): one bulk, bandwidth-efficient pass over the batch's unique remote rows (each is needed ~2.5x by
// the layer-1 kernel), run on the sampling stream one batch ahead of the training step.
void launch_stage_remote_rows(RowTable t, unsigned local_mask, const int64_t* nodes, const int32_t* cum, int n_idx,
                              int cap_nodes, void* xcache, cudaStream_t s);

// ---- sage.cu (GraphSAGE engine kernels) ---------------------------------------
struct SageAggArgs {
  // source rows: either a RowTable indexed by global id (layer 1) or a dense
  // local matrix (deeper layers, src_local != nullptr).
  RowTable feat;
  const int64_t* nodes;         // local -> global id (layer 1)
  const void* src_local;        // bf16 [n_src, d] (deeper layers)
  int d;                        // feature width (elements, multiple of 8)
  const int32_t* cum;           // device node counts
  int n_hops_targets;           // targets = cum[n_hops_targets]
  int cap_targets;
  // ELL blocks per hop (targets of hop h use ell[h] with stride k[h])
  const int32_t* ell[4];
  int k[4];
  const int32_t* deg;
  void* out;                    // bf16 [cap_targets, 2d] = [mean | self]
  // column-block output (heterogeneous layers: A_t = [mean_rel1 | mean_rel2 | ... | self]); out_ld == 0 keeps
  // the homogeneous layout above (out_ld = 2d, mean_col = 0, self_col = d)
  int out_ld;                   // row pitch of `out` in elements
  int mean_col;                 // first column of this relation's mean block
  int self_col;                 // first column of the self block, -1 = do not write it
};
void launch_sage_aggregate(const SageAggArgs& a, cudaStream_t s);

struct SageScatterArgs {
  const void* dA;               // bf16 [cap_targets, 2d]
  int d;
  const int32_t* cum;
  int n_hops_targets;
  int cap_targets;
  const int32_t* ell[4];
  int k[4];
  const int32_t* deg;
  float* dH;                    // fp32 [cap_src, d], pre-zeroed
  int dA_ld, mean_col, self_col;  // column-block input like SageAggArgs (dA_ld == 0: homogeneous [mean | self])
};
void launch_sage_scatter_bwd(const SageScatterArgs& a, cudaStream_t s);
// dH[t, :] += dA[t, col : col + d] for t < cum[n_hops] (self block of a heterogeneous layer)
void launch_add_block_f32(const void* dA, int dA_ld, int col, int d, const int32_t* cum, int n_hops, int cap,
                          float* dH, cudaStream_t s);

// ---- transpose.cu: per-batch transposed adjacency + atomics-free backward (EXPERIMENTAL) ----------
// The forward ELL blocks are keyed by TARGET (row t lists its sampled sources).  The backward of the
// mean aggregation needs the opposite view (for a source s: all targets that sampled it).  It is
// built once per batch on the sampling stream as a CSR over local source ids, ordered by hop inside
// every segment, so that one structure serves every layer (layer l uses the hops 0..L-l prefix).
struct TransposeArgs {
  const int32_t* cum;        // device counters (cum[h] = nodes before hop h's new nodes)
  const int32_t* deg;        // [cap_nodes] valid ELL entries per target
  const int32_t* ell[4];
  int k[4];
  int cap_rows[4];
  int n_hops;                // hops 0..n_hops-1 are transposed
  int cap_nodes;             // capacity of the local id space
  int cap_edges;             // capacity of tgt
  int32_t* cnt;              // [n_hops][cap_nodes]: in-counts per hop, then cumulative over hops
  int32_t* off;              // [cap_nodes + 1]: segment start per source
  int32_t* cursor;           // [n_hops][cap_nodes]: fill cursor per hop
  int32_t* tgt;              // [cap_edges]: target local ids
  int32_t* block_sums;       // [>= cap_nodes / 1024 + 2]
  int32_t* meta;             // [cap_nodes][4] = {segment start, first in-neighbour, second in-neighbour, total in-edges}
  float* meta_inv;           // [cap_nodes][2] = 1/deg of those two (the backward gather reads one record per source
                             // instead of walking off -> tgt -> deg)
};
void launch_build_transpose(const TransposeArgs& a, cudaStream_t s);

struct SageGatherBwdArgs {
  const void* dA;            // bf16 [cap_targets, 2d] = gradient of [mean | self]
  int d;
  const int32_t* cum;
  int n_hops_targets;        // targets = cum[n_hops_targets], sources = cum[n_hops_targets + 1]
  int cap_targets;
  int cap_src;               // rows of dPre / Z
  const int32_t* deg;
  const int32_t* off;        // transposed CSR
  const int32_t* cnt_upto;   // [cap_nodes] in-edges of the hops used by this layer (prefix of the segment)
  const int32_t* tgt;
  const int32_t* meta;       // TransposeArgs::meta / meta_inv
  const float* meta_inv;
  const void* Z;             // bf16 [cap_src, d] activations of the previous layer (ReLU mask) or nullptr
  void* dPre;                // bf16 [cap_src, d]
  float* colsum;             // optional fp32 [d]: bias gradient (16-byte aligned: accumulated with red.v4)
  int colsum_prezeroed;      // the caller already zeroed colsum (no memset node in the launch chain)
  float gscale;              // 1/(1-p) when Z is a post-dropout activation, else 1
};
// dPre[s] = relu'(Z[s]) * (dA_self[s] + sum_{t in in(s)} dA_mean[t] / deg[t]); rows >= sources are zero-filled.
// Replaces zero_rows + sage_scatter_bwd (fp32 atomics) + relu_bwd_cast.
void launch_sage_gather_bwd(const SageGatherBwdArgs& a, cudaStream_t s);

// dPre = (Z > 0) ? bf16(dH) : 0 for rows < cum[n_hops]; 0 beyond.
// colsum (optional, fp32 [d]): fused bias gradient = column sums of dPre.
void launch_relu_bwd_cast(const float* dH, const void* Z, const int32_t* cum, int n_hops, int cap,
                          int d, void* dPre, float* colsum, cudaStream_t s, bool prezeroed = false,
                          float gscale = 1.f);
// In-place inverted dropout (keep prob 1-p, kept values scaled by 1/(1-p)) on rows < cum[n_hops] of bf16 Z[cap, d];
// mask = Philox(seed, layer, *step_dev, element).  Backward: launch_relu_bwd_cast(..., gscale = 1/(1-p)) on the
// post-dropout Z.
void launch_dropout_bf16(void* Z, const int32_t* cum, int n_hops, int cap, int d, float p, uint64_t seed, int layer,
                         const int32_t* step_dev, cudaStream_t s);
// g[0..n) = 0, *loss = 0, *correct = 0 (one launch at the start of the gradient phase; the *_prezeroed variants of
// the kernels below then skip their own memsets, which keeps the step a pure kernel chain)
void launch_zero_grads(float* g, int64_t n, float* loss, int32_t* correct, cudaStream_t s);
// bias + relu epilogue over valid rows (zero beyond): Z = relu(Z + b)
void launch_bias_relu(void* Z, const void* bias, const int32_t* cum, int n_hops, int cap, int d,
                      int relu, cudaStream_t s);
// loss = mean NLL(log_softmax(logits[:n0, :C]), y); dlogits = (softmax - onehot)/n0
// labels come from y[r], or (labels_all != nullptr) from labels_all[nodes[r]].
void launch_softmax_nll(const void* logits, int ld, int C, const int64_t* y, const int64_t* labels_all,
                        const int64_t* nodes, const int32_t* cum, int cap, float* loss, void* dlogits,
                        int32_t* correct, float* colsum, cudaStream_t s, bool prezeroed = false);
// out[c] = sum over valid rows of X[:, c]  (d multiple of 8, d <= 2048)
void launch_colsum_bf16(const void* X, const int32_t* cum, int n_hops, int cap, int d, float* out,
                        cudaStream_t s);
void launch_zero_rows(float* p, const int32_t* cum, int n_hops, int cap, int d, cudaStream_t s);
// flat fp32 Adam over [n] with bf16 shadow copy refresh.
void launch_adam(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr,
                 float b1, float b2, float eps, float wd, int32_t* step_dev, float gscale,
                 cudaStream_t s);
void launch_bf16_to_f32(const void* src, float* dst, int64_t n, cudaStream_t s);

// ---- peer.cu (collectives over NVLink peer memory, CUDA-graph capturable) ------
struct PeerPtrs {
  int world, rank;
  const float* g[kMaxParts];   // gradient buffer of every rank, mapped on this device
  int32_t* flags[kMaxParts];   // barrier flag array [2 * world] of every rank, mapped on this device
};
void launch_peer_barrier(const PeerPtrs& p, int which, int32_t* epoch_dev, int32_t* err, cudaStream_t s);
void launch_adam_peer(const PeerPtrs& p, float* param, float* m, float* v, void* p_bf16, int64_t n, float lr,
                      float b1, float b2, float eps, float wd, int32_t* step_dev, float gscale,
                      cudaStream_t s, const int32_t* err = nullptr);

// ---- sage_tc.cu (tcgen05 fused gather+aggregate+GEMM) -------------------------
struct SageFusedArgs {
  SageAggArgs agg;              // agg.out may be nullptr (A tile is not materialised)
  const void* w_packed;         // bf16, UMMA SWIZZLE_128B K-major image, [K/64][N][64]
  const void* bias;             // bf16 [N]
  int n_out;                    // N (multiple of 16, <= 256)
  int relu;
  void* z;                      // bf16 [cap_targets, N]
  void* a_save;                 // optional bf16 [cap_targets, 2d] for backward
  unsigned long long* trace;    // optional per-CTA clock64 timeline (diagnostics, see sage_fused_trace)
  int feat_fp8;                 // 1: agg.feat rows are MXFP8 (d e4m3 bytes + d/32 UE8M0 scales, 16-byte padded)
  int l2_prefetch;              // resolver warps prefetch next tile's local feature rows into L2
  unsigned local_mask;          // bit p set: part p of agg.feat lives in this GPU's HBM (prefetchable into its L2)
  // Staged remote rows: rows of the batch's nodes that live in PEER HBM, copied once per batch (on the sampling
  // stream, launch_stage_remote_rows) into a local buffer indexed by the node's LOCAL id.  nullptr = read peers in
  // place.  cache_part: index of a non-local part whose handle slot is reused for the cache.
  const void* xcache;
  int cache_part;
};
int sage_fused_supported(int d, int n_out);
// Copies the per-CTA timeline of the last traced launch (GLT_B200_FUSED_TRACE=1) to `host` [148*32].
void sage_fused_trace_copy(unsigned long long* host);
constexpr int kFusedTraceSlots = 32;
void launch_sage_fused(const SageFusedArgs& a, int num_sms, cudaStream_t s);
// W [N, K] row-major bf16 -> packed swizzled image used by launch_sage_fused.
void launch_pack_weight(const void* w, int n, int k, void* packed, cudaStream_t s);


// ---- tc_gemm.cu (TMA-fed tcgen05 GEMMs: forward layers 2..L, dA, split-K dW) ---------------------
struct TcProblem {
  int a_mn, b_mn;          // operand major-ness: 0 = K-major (row-major [rows, K]), 1 = MN-major ([K rows, M/N cols])
  int epi;                 // 0: bf16 out (+bias, +ReLU) through a TMA store; 1: fp32 red-add (split-K)
  int bn;                  // tile width: 64 / 128 / 256, divides n
  int m, n, k;             // static extents (capacity for the dynamic one)
  const int32_t* dyn;      // device counters; the dynamic extent is min(dyn[dyn_idx], dyn_cap)
  int dyn_idx, dyn_cap;
  int dyn_is_k;            // 0: the dynamic extent is M (forward / dA); 1: it is K (dW)
  const void* bias;        // bf16 [n] or nullptr (epi 0)
  int relu;
  float* out32;            // epi 1: fp32 [m_valid, ld32]
  int ld32, m_valid;
};
struct TcGemmArgs {
  int n_prob;
  TcProblem p[2];
};
struct alignas(64) TcGemmLaunch {
  unsigned char maps[6][128];   // CUtensorMap A0, B0, C0, A1, B1, C1
  TcGemmArgs args;
  int max_items;                // upper bound of work items (grid = min(#SMs, max_items)); 0 = #SMs
};
// bf16 [rows, cols] row-major (pitch ld elements) tensor map with a {box_cols = 64, box_rows} SWIZZLE_128B box.
// Returns 0 on success (1: driver entry point unavailable, 2: encode failed).
int make_tmap_bf16_2d(void* out_map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                      int box_rows);
// uint8 (e4m3) [rows, cols] row-major tensor map with a {box_cols = 128 bytes, box_rows} SWIZZLE_128B box
int make_tmap_u8_2d(void* out_map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                    int box_rows);
size_t tc_gemm_smem_bytes();

// ---- tc_gemm_mx.cu (block-scaled MXFP8 forward GEMM, tcgen05.mma.kind::mxf8f6f4.block_scale) -----------------
struct TcMxArgs {
  int m, n, k;              // rows of A (capacity), rows of W, contraction length (multiple of 128)
  int bn;                   // 128 or 256, divides n
  const int32_t* dyn;       // optional device counter for the valid rows of A
  int dyn_idx;
  const void* sfa;          // packed UE8M0 scale blocks of A: [ceil(m/128)][k/128][512 B]
  const void* sfb;          // packed UE8M0 scale blocks of W: [n/128][k/128][512 B]
  const void* bias;         // bf16 [n] or nullptr
  int relu;
};
void launch_tc_gemm_mx(const void* maps3, const TcMxArgs& a, int num_sms, cudaStream_t s);
// programmatic-dependent-launch switch (launch_utils.h); returns the previous value
int set_pdl(int on);
void launch_tc_gemm(const TcGemmLaunch& L, int num_sms, cudaStream_t s);

}  // namespace glt
