// Python bindings: the only translation unit (besides cpu/*.cc) that sees torch
// headers.  Native surface parity with the reference's pybind module
// (python/py_export_glt.cc:47-222): Graph, samplers, negative sampler, inducer
// tables, subgraph op, SampleQueue, UnifiedTensor-style row tables -- plus the
// batch sampler arena and the GraphSAGE engine kernels that have no reference
// counterpart.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <pybind11/functional.h>
#include <pybind11/stl.h>
#include <torch/extension.h>

#include "cpu/cpu_ops.h"
#include "cpu/sample_queue.h"
#include "cuda/glt_cuda.h"

namespace py = pybind11;
using torch::Tensor;

namespace glt {

static cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }

static void check_cuda_err(const char* what) {
  cudaError_t e = cudaGetLastError();
  TORCH_CHECK(e == cudaSuccess, what, ": ", cudaGetErrorString(e));
}

static const void* dev_ptr(const Tensor& t) {
  // CUDA tensors (local or peer device) and pinned host tensors are both
  // directly dereferenceable from a kernel under UVA.
  TORCH_CHECK(t.is_cuda() || t.is_pinned(), "tensor must be CUDA or pinned host memory");
  return t.data_ptr();
}

// ---------------------------------------------------------------------------
// GraphHandle: device view of a (possibly multi-GPU) CSR.
// ---------------------------------------------------------------------------
struct GraphHandle {
  GraphTable tbl{};
  std::vector<Tensor> keep;
  int device = 0;
  bool has_eids = true, has_weights = true;
  int64_t num_rows = 0;

  explicit GraphHandle(int dev) : device(dev) { tbl.num_parts = 0; tbl.idx_bytes = 8; }

  void add_shard(const Tensor& indptr, const Tensor& indices, const c10::optional<Tensor>& eids,
                 const c10::optional<Tensor>& weights, int64_t row_begin, int64_t row_end) {
    TORCH_CHECK(tbl.num_parts < kMaxParts, "too many graph shards");
    TORCH_CHECK(indptr.scalar_type() == torch::kInt64 && indptr.is_contiguous());
    TORCH_CHECK(indices.is_contiguous());
    int ib = indices.scalar_type() == torch::kInt32 ? 4 : 8;
    TORCH_CHECK(indices.scalar_type() == torch::kInt32 || indices.scalar_type() == torch::kInt64);
    if (tbl.num_parts == 0) tbl.idx_bytes = ib;
    TORCH_CHECK(tbl.idx_bytes == ib, "all shards must share the column id width");
    TORCH_CHECK(indptr.numel() == row_end - row_begin + 1, "indptr size != rows + 1");
    CsrShard& s = tbl.parts[tbl.num_parts++];
    s.indptr = reinterpret_cast<const int64_t*>(dev_ptr(indptr));
    s.indices = dev_ptr(indices);
    s.eids = nullptr;
    s.weights = nullptr;
    keep.push_back(indptr);
    keep.push_back(indices);
    if (eids.has_value() && eids->defined()) {
      TORCH_CHECK(eids->scalar_type() == torch::kInt64);
      s.eids = reinterpret_cast<const int64_t*>(dev_ptr(*eids));
      keep.push_back(*eids);
    } else {
      has_eids = false;
    }
    if (weights.has_value() && weights->defined()) {
      TORCH_CHECK(weights->scalar_type() == torch::kFloat32);
      s.weights = reinterpret_cast<const float*>(dev_ptr(*weights));
      keep.push_back(*weights);
    } else {
      has_weights = false;
    }
    s.row_begin = row_begin;
    s.row_end = row_end;
    num_rows = std::max(num_rows, row_end);
  }

  torch::TensorOptions opts(torch::ScalarType t) const {
    return torch::TensorOptions().dtype(t).device(torch::kCUDA, device);
  }

  // (nbrs [n,k] padded with -1, counts [n] int32, eids [n,k] | undefined)
  std::tuple<Tensor, Tensor, Tensor> sample_one_hop(const Tensor& seeds, int64_t k, bool with_edge,
                                                    bool weighted, bool replace, int64_t seed,
                                                    int64_t stream) {
    c10::cuda::CUDAGuard guard(device);
    TORCH_CHECK(seeds.is_cuda() && seeds.scalar_type() == torch::kInt64 && seeds.is_contiguous());
    TORCH_CHECK(k > 0 && k <= 512, "fanout must be in [1, 512] (use -1 for all neighbours)");
    TORCH_CHECK(!with_edge || has_eids, "graph has no edge ids");
    TORCH_CHECK(!weighted || has_weights, "graph has no edge weights on the device");
    const int64_t n = seeds.numel();
    Tensor nbrs = torch::empty({n, k}, opts(torch::kInt64));
    Tensor cnt = torch::empty({n}, opts(torch::kInt32));
    Tensor eids = with_edge ? torch::empty({n, k}, opts(torch::kInt64)) : Tensor();
    launch_sample_one_hop(tbl, seeds.data_ptr<int64_t>(), n, k, weighted, replace, seed, stream,
                          nbrs.data_ptr<int64_t>(), with_edge ? eids.data_ptr<int64_t>() : nullptr,
                          cnt.data_ptr<int32_t>(), cur_stream());
    check_cuda_err("sample_one_hop");
    return {nbrs, cnt, eids};
  }

  Tensor lookup_degree(const Tensor& ids) {
    c10::cuda::CUDAGuard guard(device);
    Tensor out = torch::empty({ids.numel()}, opts(torch::kInt64));
    launch_lookup_degree(tbl, ids.data_ptr<int64_t>(), ids.numel(), out.data_ptr<int64_t>(), cur_stream());
    check_cuda_err("lookup_degree");
    return out;
  }

  // all neighbours; returns (nbrs [sum], counts [n] int64, eids)
  std::tuple<Tensor, Tensor, Tensor> full_neighbors(const Tensor& ids, bool with_edge) {
    c10::cuda::CUDAGuard guard(device);
    Tensor deg = lookup_degree(ids);
    Tensor offs = torch::zeros({ids.numel() + 1}, opts(torch::kInt64));
    if (ids.numel() > 0) offs.narrow(0, 1, ids.numel()).copy_(deg.cumsum(0));
    const int64_t total = ids.numel() > 0 ? offs[ids.numel()].item<int64_t>() : 0;  // host sync
    Tensor nbrs = torch::empty({total}, opts(torch::kInt64));
    Tensor eids = with_edge ? torch::empty({total}, opts(torch::kInt64)) : Tensor();
    launch_copy_neighbors(tbl, ids.data_ptr<int64_t>(), ids.numel(), offs.data_ptr<int64_t>(),
                          nbrs.data_ptr<int64_t>(), with_edge ? eids.data_ptr<int64_t>() : nullptr,
                          cur_stream());
    check_cuda_err("full_neighbors");
    return {nbrs, deg, eids};
  }

  // (rows, cols, count) with rows/cols sized `req`; count stays on the device
  std::tuple<Tensor, Tensor, Tensor> negative_sample(int64_t num_rows_, int64_t num_cols, int64_t req,
                                                     int64_t trials, bool padding, int64_t seed,
                                                     int64_t stream) {
    c10::cuda::CUDAGuard guard(device);
    Tensor rows = torch::empty({req}, opts(torch::kInt64));
    Tensor cols = torch::empty({req}, opts(torch::kInt64));
    Tensor count = torch::zeros({1}, opts(torch::kInt32));
    launch_negative_sample(tbl, num_rows_, num_cols, req, trials, padding, seed, stream,
                           rows.data_ptr<int64_t>(), cols.data_ptr<int64_t>(),
                           count.data_ptr<int32_t>(), cur_stream());
    check_cuda_err("negative_sample");
    return {rows, cols, count};
  }

  Tensor random_walk(const Tensor& starts, int64_t walk_len, double p, double q, int64_t seed,
                     int64_t stream) {
    c10::cuda::CUDAGuard guard(device);
    Tensor out = torch::empty({starts.numel(), walk_len + 1}, opts(torch::kInt64));
    launch_random_walk(tbl, starts.data_ptr<int64_t>(), starts.numel(), walk_len, p, q, seed, stream,
                       out.data_ptr<int64_t>(), cur_stream());
    check_cuda_err("random_walk");
    return out;
  }

  Tensor nbr_prob(GraphHandle& nbr_graph, const Tensor& last, const Tensor& nbr_last, int64_t k) {
    c10::cuda::CUDAGuard guard(device);
    Tensor cur = torch::zeros_like(last);
    launch_nbr_prob(tbl, nbr_graph.tbl, last.data_ptr<float>(), nbr_last.data_ptr<float>(),
                    std::min<int64_t>(last.numel(), num_rows), nbr_last.numel(), k,
                    cur.data_ptr<float>(), cur_stream());
    check_cuda_err("nbr_prob");
    return cur;
  }
};

// ---------------------------------------------------------------------------
// DeviceTable: GPU id table (inducer state); one per node type for hetero.
// ---------------------------------------------------------------------------
struct DeviceTable {
  int device;
  int64_t cap_nodes;
  Tensor keys, vals, aux, nodes, cursor;
  HashTable ht{};

  DeviceTable(int dev, int64_t capacity) : device(dev), cap_nodes(std::max<int64_t>(capacity, 16)) {
    c10::cuda::CUDAGuard guard(device);
    int64_t slots = 64;
    while (slots < cap_nodes * 2) slots <<= 1;
    auto o64 = torch::TensorOptions().dtype(torch::kInt64).device(torch::kCUDA, device);
    auto o32 = torch::TensorOptions().dtype(torch::kInt32).device(torch::kCUDA, device);
    keys = torch::full({slots}, -1, o64);
    vals = torch::zeros({slots}, o32);
    aux = torch::zeros({slots}, o32);
    nodes = torch::zeros({cap_nodes}, o64);
    cursor = torch::zeros({1}, o32);
    ht.keys = keys.data_ptr<int64_t>();
    ht.vals = vals.data_ptr<int32_t>();
    ht.aux = aux.data_ptr<int32_t>();
    ht.mask = static_cast<uint32_t>(slots - 1);
  }

  void clear() {
    c10::cuda::CUDAGuard guard(device);
    launch_table_clear(ht, cur_stream());
    cursor.zero_();
  }

  // unordered insert; returns local ids (int32).  New keys are appended to `nodes`.
  Tensor insert(const Tensor& k) {
    c10::cuda::CUDAGuard guard(device);
    TORCH_CHECK(k.is_cuda() && k.scalar_type() == torch::kInt64 && k.is_contiguous());
    Tensor out = torch::empty({k.numel()}, torch::TensorOptions().dtype(torch::kInt32).device(k.device()));
    launch_table_insert(ht, k.data_ptr<int64_t>(), k.numel(), nodes.data_ptr<int64_t>(),
                        cursor.data_ptr<int32_t>(), cap_nodes, out.data_ptr<int32_t>(), cur_stream());
    launch_table_resolve(ht, out.data_ptr<int32_t>(), k.numel(), cur_stream());
    check_cuda_err("table insert");
    return out;
  }

  // ordered (first-occurrence) insert into an empty table; returns local ids
  Tensor init_ordered(const Tensor& k) {
    c10::cuda::CUDAGuard guard(device);
    TORCH_CHECK(k.is_cuda() && k.scalar_type() == torch::kInt64 && k.is_contiguous());
    TORCH_CHECK(k.numel() <= cap_nodes, "table capacity exceeded");
    auto o32 = torch::TensorOptions().dtype(torch::kInt32).device(k.device());
    Tensor out = torch::empty({k.numel()}, o32);
    Tensor scratch = torch::empty({std::max<int64_t>(k.numel(), 1)}, o32);
    Tensor counters = torch::zeros({16}, o32);
    BatchCounters c{counters.data_ptr<int32_t>(), counters.data_ptr<int32_t>() + 6, cursor.data_ptr<int32_t>(), nullptr};
    launch_init_seeds(k.data_ptr<int64_t>(), k.numel(), nullptr, ht, nodes.data_ptr<int64_t>(),
                      out.data_ptr<int32_t>(), scratch.data_ptr<int32_t>(), c, nullptr, 0, cur_stream());
    check_cuda_err("table init_ordered");
    return out;
  }

  Tensor lookup(const Tensor& k) {
    c10::cuda::CUDAGuard guard(device);
    Tensor out = torch::empty({k.numel()}, torch::TensorOptions().dtype(torch::kInt32).device(k.device()));
    launch_table_lookup(ht, k.data_ptr<int64_t>(), k.numel(), out.data_ptr<int32_t>(), cur_stream());
    check_cuda_err("table lookup");
    return out;
  }

  int64_t size() { return cursor.item<int32_t>(); }  // host sync

  // induced subgraph of `g` on the table's current node set (n = size())
  std::tuple<Tensor, Tensor, Tensor> subgraph(GraphHandle& g, int64_t n, bool with_edge) {
    c10::cuda::CUDAGuard guard(device);
    auto o64 = torch::TensorOptions().dtype(torch::kInt64).device(torch::kCUDA, device);
    Tensor cnt = torch::zeros({n}, o64);
    launch_subgraph_count(g.tbl, ht, nodes.data_ptr<int64_t>(), n, cnt.data_ptr<int64_t>(), cur_stream());
    Tensor offs = torch::zeros({n + 1}, o64);
    if (n > 0) offs.narrow(0, 1, n).copy_(cnt.cumsum(0));
    const int64_t total = n > 0 ? offs[n].item<int64_t>() : 0;  // host sync
    Tensor rows = torch::empty({total}, o64), cols = torch::empty({total}, o64);
    Tensor eids = with_edge ? torch::empty({total}, o64) : Tensor();
    launch_subgraph_fill(g.tbl, ht, nodes.data_ptr<int64_t>(), n, offs.data_ptr<int64_t>(),
                         rows.data_ptr<int64_t>(), cols.data_ptr<int64_t>(),
                         with_edge ? eids.data_ptr<int64_t>() : nullptr, cur_stream());
    check_cuda_err("subgraph");
    return {rows, cols, eids};
  }
};

// ---------------------------------------------------------------------------
// SamplerArena: static-shape multi-hop sampling state (no host sync).
// ---------------------------------------------------------------------------
struct SamplerArena {
  int device;
  int max_seeds;
  std::vector<int64_t> fanouts;
  bool with_edge;
  std::vector<int64_t> cap_rows;  // frontier capacity per hop
  int64_t cap_nodes;
  Tensor nodes, deg, counters, seed_local, scratch, step;
  std::vector<Tensor> ell, ell_eids;
  std::unique_ptr<DeviceTable> table;
  // optional transposed adjacency (EXPERIMENTAL, see cuda/transpose.cu): built after the hops of
  // every sample() once enable_transpose(n_hops) was called
  int tr_hops = 0;
  int64_t tr_cap_edges = 0;
  Tensor tr_cnt, tr_off, tr_cursor, tr_tgt, tr_block_sums, tr_meta, tr_meta_inv;

  // cap_override: optional calibrated frontier capacities per hop (+ one trailing entry for the
  // number of nodes the last hop may add); empty = worst case (max_seeds * prod(fanouts)).
  SamplerArena(int dev, int64_t max_seeds_, std::vector<int64_t> fanouts_, bool with_edge_,
               int64_t num_graph_nodes, std::vector<int64_t> cap_override = {})
      : device(dev), max_seeds(max_seeds_), fanouts(std::move(fanouts_)), with_edge(with_edge_) {
    c10::cuda::CUDAGuard guard(device);
    TORCH_CHECK(fanouts.size() >= 1 && fanouts.size() <= 4, "1..4 hops supported by the arena");
    auto o64 = torch::TensorOptions().dtype(torch::kInt64).device(torch::kCUDA, device);
    auto o32 = torch::TensorOptions().dtype(torch::kInt32).device(torch::kCUDA, device);
    int64_t rows = max_seeds;
    cap_nodes = max_seeds;
    TORCH_CHECK(cap_override.empty() || cap_override.size() == fanouts.size() + 1,
                "cap_override needs len(fanouts)+1 entries");
    for (size_t h = 0; h < fanouts.size(); ++h) {
      TORCH_CHECK(fanouts[h] > 0 && fanouts[h] <= 512, "arena fanouts must be in [1,512]");
      if (!cap_override.empty()) rows = std::min(rows, std::max<int64_t>(cap_override[h], 1));
      cap_rows.push_back(rows);
      ell.push_back(torch::full({rows * fanouts[h]}, -1, o32));
      ell_eids.push_back(with_edge ? torch::full({rows * fanouts[h]}, -1, o64) : Tensor());
      rows = std::min<int64_t>(rows * fanouts[h], num_graph_nodes > 0 ? num_graph_nodes : INT64_MAX);
      if (!cap_override.empty()) rows = std::min(rows, std::max<int64_t>(cap_override[h + 1], 1));
      cap_nodes += rows;
    }
    cap_rows.push_back(rows);  // capacity of the nodes added by the last hop
    if (num_graph_nodes > 0) cap_nodes = std::min(cap_nodes, num_graph_nodes + max_seeds);
    TORCH_CHECK(cap_nodes < (1LL << 30), "sampler arena too large");
    table = std::make_unique<DeviceTable>(device, cap_nodes);
    nodes = table->nodes;
    deg = torch::zeros({cap_nodes}, o32);
    counters = torch::zeros({16}, o32);
    seed_local = torch::zeros({max_seeds}, o32);
    scratch = torch::zeros({max_seeds}, o32);
    step = torch::zeros({1}, o32);
  }

  void enable_transpose(int64_t n_hops) {
    c10::cuda::CUDAGuard guard(device);
    TORCH_CHECK(n_hops >= 1 && n_hops <= static_cast<int64_t>(fanouts.size()), "bad number of hops to transpose");
    auto o32 = torch::TensorOptions().dtype(torch::kInt32).device(torch::kCUDA, device);
    tr_hops = static_cast<int>(n_hops);
    tr_cap_edges = 0;
    for (int h = 0; h < tr_hops; ++h) tr_cap_edges += cap_rows[h] * fanouts[h];
    TORCH_CHECK(tr_cap_edges < (1LL << 31), "transposed adjacency too large");
    tr_cnt = torch::zeros({tr_hops, cap_nodes}, o32);
    tr_cursor = torch::zeros({tr_hops, cap_nodes}, o32);
    tr_off = torch::zeros({cap_nodes + 1}, o32);
    tr_tgt = torch::zeros({std::max<int64_t>(tr_cap_edges, 1)}, o32);
    tr_block_sums = torch::zeros({cap_nodes / 1024 + 2}, o32);
    tr_meta = torch::zeros({cap_nodes, 4}, o32);
    tr_meta_inv = torch::zeros({cap_nodes, 2}, o32.dtype(torch::kFloat32));
  }

  void build_transpose(cudaStream_t s) {
    TransposeArgs a{};
    a.cum = counters.data_ptr<int32_t>();
    a.deg = deg.data_ptr<int32_t>();
    for (int h = 0; h < 4; ++h) { a.ell[h] = nullptr; a.k[h] = 1; a.cap_rows[h] = 0; }
    for (int h = 0; h < tr_hops; ++h) {
      a.ell[h] = ell[h].data_ptr<int32_t>();
      a.k[h] = static_cast<int>(fanouts[h]);
      a.cap_rows[h] = static_cast<int>(cap_rows[h]);
    }
    a.n_hops = tr_hops;
    a.cap_nodes = static_cast<int>(cap_nodes);
    a.cap_edges = static_cast<int>(tr_cap_edges);
    a.cnt = tr_cnt.data_ptr<int32_t>();
    a.off = tr_off.data_ptr<int32_t>();
    a.cursor = tr_cursor.data_ptr<int32_t>();
    a.tgt = tr_tgt.data_ptr<int32_t>();
    a.block_sums = tr_block_sums.data_ptr<int32_t>();
    a.meta = tr_meta.data_ptr<int32_t>();
    a.meta_inv = tr_meta_inv.data_ptr<float>();
    launch_build_transpose(a, s);
  }

  BatchCounters bc() {
    int32_t* c = counters.data_ptr<int32_t>();
    return BatchCounters{c, c + 6, table->cursor.data_ptr<int32_t>(), c + 12};
  }

  bool deterministic = false;   // re-assign every hop's new local ids in ascending global-id order
  Tensor det_tmp;

  void sample(GraphHandle& g, const Tensor& seeds, const c10::optional<Tensor>& n_dev, int64_t seed,
              int64_t stream_base, bool weighted, bool replace, bool use_dev_step, int64_t step_inc) {
    c10::cuda::CUDAGuard guard(device);
    TORCH_CHECK(seeds.is_cuda() && seeds.scalar_type() == torch::kInt64 && seeds.is_contiguous());
    TORCH_CHECK(seeds.numel() <= max_seeds, "more seeds than the arena was built for");
    TORCH_CHECK(!with_edge || g.has_eids, "graph has no edge ids");
    TORCH_CHECK(!weighted || g.has_weights, "graph has no device edge weights");
    cudaStream_t s = cur_stream();
    launch_table_clear(table->ht, s);
    const int32_t* nd = (n_dev.has_value() && n_dev->defined()) ? n_dev->data_ptr<int32_t>() : nullptr;
    launch_init_seeds(seeds.data_ptr<int64_t>(), seeds.numel(), nd, table->ht, nodes.data_ptr<int64_t>(),
                      seed_local.data_ptr<int32_t>(), scratch.data_ptr<int32_t>(), bc(),
                      (use_dev_step && step_inc != 0) ? step.data_ptr<int32_t>() : nullptr,
                      static_cast<int>(step_inc), s);
    for (size_t h = 0; h < fanouts.size(); ++h) {
      HopArgs a{};
      a.g = g.tbl;
      a.t = table->ht;
      a.c = bc();
      a.nodes = nodes.data_ptr<int64_t>();
      a.ell = ell[h].data_ptr<int32_t>();
      a.ell_eids = with_edge ? ell_eids[h].data_ptr<int64_t>() : nullptr;
      a.deg = deg.data_ptr<int32_t>();
      a.hop = h;
      a.k = fanouts[h];
      a.cap_rows = cap_rows[h];
      a.cap_nodes = cap_nodes;
      a.cap_rows_next = cap_rows[h + 1];
      a.weighted = weighted;
      a.replace = replace;
      a.seed = seed;
      a.stream = static_cast<uint32_t>(stream_base + h);
      a.stream_dev = use_dev_step ? step.data_ptr<int32_t>() : nullptr;
      launch_sample_hop(a, s);
      if (deterministic) {
        if (!det_tmp.defined())
          det_tmp = torch::empty({cap_nodes}, torch::TensorOptions().dtype(torch::kInt64).device(torch::kCUDA, device));
        launch_det_keys(a, det_tmp.data_ptr<int64_t>(), s);
        Tensor sorted = std::get<0>(torch::sort(det_tmp));
        launch_det_assign(a, sorted.data_ptr<int64_t>(), s);
      }
      launch_relabel_hop(a, s);
    }
    if (tr_hops > 0) build_transpose(s);
    check_cuda_err("arena sample");
  }

  // PyG-shaped COO (one host sync to learn the sizes):
  // returns (node [N], row [E], col [E], eids|undef, num_sampled_nodes, num_sampled_edges)
  std::tuple<Tensor, Tensor, Tensor, Tensor, std::vector<int64_t>, std::vector<int64_t>> to_coo() {
    c10::cuda::CUDAGuard guard(device);
    auto o64 = torch::TensorOptions().dtype(torch::kInt64).device(torch::kCUDA, device);
    Tensor host = counters.cpu();  // sync
    const int32_t* c = host.data_ptr<int32_t>();
    const int L = fanouts.size();
    std::vector<int64_t> nn, ne;
    for (int h = 0; h <= L; ++h) nn.push_back(c[h + 1] - c[h]);
    int64_t E = 0;
    for (int h = 0; h < L; ++h) { ne.push_back(c[6 + h]); E += c[6 + h]; }
    const int64_t N = c[L + 1], T = c[L];
    Tensor offs = torch::zeros({T + 1}, o64);
    if (T > 0) offs.narrow(0, 1, T).copy_(deg.narrow(0, 0, T).to(torch::kInt64).cumsum(0));
    Tensor rows = torch::empty({E}, o64), cols = torch::empty({E}, o64);
    Tensor eids = with_edge ? torch::empty({E}, o64) : Tensor();
    for (int h = 0; h < L; ++h) {
      launch_ell_to_coo(ell[h].data_ptr<int32_t>(), with_edge ? ell_eids[h].data_ptr<int64_t>() : nullptr,
                        deg.data_ptr<int32_t>(), offs.data_ptr<int64_t>(), counters.data_ptr<int32_t>(),
                        h, fanouts[h], cap_rows[h], rows.data_ptr<int64_t>(), cols.data_ptr<int64_t>(),
                        with_edge ? eids.data_ptr<int64_t>() : nullptr, cur_stream());
    }
    check_cuda_err("to_coo");
    return {nodes.narrow(0, 0, N).clone(), rows, cols, eids, nn, ne};
  }
};


// ---------------------------------------------------------------------------
// HeteroArena: device-resident multi-hop sampling + inducing over a heterogeneous graph (native counterpart of
// the reference's CUDAHeteroInducer, csrc/cuda/inducer.cu:194-338, plus the per-hop orchestration of
// python/sampler/neighbor_sampler.py:232-317).  Node types and relations are integer coded; every node type
// owns an id table / id list / hop counters, every relation owns fixed-stride ELL blocks per hop.  One hop =
// three launches for ALL relations together (grouped sample, per-type finalize, grouped relabel) and no host
// synchronisation; sizes stay in `counters` until a PyG-shaped COO is requested.
// ---------------------------------------------------------------------------
struct HeteroArena {
  int device, n_types, n_rel, hops;
  bool with_edge;
  std::vector<GraphHandle*> graphs;
  std::vector<int64_t> key_type, nbr_type;
  std::vector<std::vector<int64_t>> fanouts;       // [rel][hop]
  std::vector<std::vector<int64_t>> cap_rows;      // [type][hop + 1]
  std::vector<int64_t> cap_nodes, max_seeds;
  std::vector<std::unique_ptr<DeviceTable>> tables;
  Tensor all_keys;                                 // the tables' key arrays, contiguous: one memset clears them
  Tensor counters, step, descs, type_states, seed_scratch, seed_local, deg_all;
  std::vector<Tensor> deg;                         // [rel] views into deg_all
  std::vector<std::vector<Tensor>> ell, ell_eids;  // [rel][hop]
  std::vector<int> max_k, max_rows;                // per hop
  bool any_zero_fanout = false;
  int ctr_type_base(int t) const { return t * 8; }
  int ctr_rel_base(int r) const { return n_types * 8 + r * 4; }
  int ctr_overflow() const { return n_types * 8 + n_rel * 4; }

  HeteroArena(int dev, int64_t n_types_, std::vector<GraphHandle*> graphs_, std::vector<int64_t> key_type_,
              std::vector<int64_t> nbr_type_, std::vector<std::vector<int64_t>> fanouts_,
              std::vector<int64_t> num_nodes, std::vector<int64_t> max_seeds_, bool with_edge_, int64_t seed,
              bool weighted, bool replace, int64_t cap_limit, std::vector<std::vector<int64_t>> cap_override)
      : device(dev), n_types(n_types_), n_rel(graphs_.size()), with_edge(with_edge_), graphs(std::move(graphs_)),
        key_type(std::move(key_type_)), nbr_type(std::move(nbr_type_)), fanouts(std::move(fanouts_)),
        max_seeds(std::move(max_seeds_)) {
    c10::cuda::CUDAGuard guard(device);
    TORCH_CHECK(n_rel >= 1 && static_cast<int>(key_type.size()) == n_rel && static_cast<int>(nbr_type.size()) == n_rel &&
                static_cast<int>(fanouts.size()) == n_rel, "HeteroArena: one (graph, key type, nbr type, fanouts) per relation");
    TORCH_CHECK(static_cast<int>(num_nodes.size()) == n_types && static_cast<int>(max_seeds.size()) == n_types);
    hops = fanouts[0].size();
    TORCH_CHECK(hops >= 1 && hops <= 4, "1..4 hops supported by the arena");
    for (int r = 0; r < n_rel; ++r) {
      TORCH_CHECK(static_cast<int>(fanouts[r].size()) == hops, "all relations need the same number of hops");
      TORCH_CHECK(key_type[r] >= 0 && key_type[r] < n_types && nbr_type[r] >= 0 && nbr_type[r] < n_types);
      TORCH_CHECK(!with_edge || graphs[r]->has_eids, "graph has no edge ids");
      TORCH_CHECK(!weighted || graphs[r]->has_weights, "graph has no device edge weights");
      for (int h = 0; h < hops; ++h) {
        TORCH_CHECK(fanouts[r][h] >= 0 && fanouts[r][h] <= 512, "arena fanouts must be in [0,512]");
        if (fanouts[r][h] == 0) any_zero_fanout = true;
      }
    }
    if (cap_limit <= 0) cap_limit = INT64_MAX;
    // frontier capacities: worst case of the fan-out recursion, bounded by the type's node count and cap_limit
    cap_rows.assign(n_types, std::vector<int64_t>(hops + 1, 0));
    for (int t = 0; t < n_types; ++t) cap_rows[t][0] = max_seeds[t];
    for (int h = 0; h < hops; ++h)
      for (int t = 0; t < n_types; ++t) {
        int64_t rows = 0;
        for (int r = 0; r < n_rel; ++r)
          if (nbr_type[r] == t) rows += cap_rows[key_type[r]][h] * fanouts[r][h];
        rows = std::min(rows, std::min(num_nodes[t] > 0 ? num_nodes[t] : INT64_MAX, cap_limit));
        if (!cap_override.empty()) rows = std::min(rows, std::max<int64_t>(cap_override[t][h + 1], 0));
        cap_rows[t][h + 1] = rows;
      }
    auto o64 = torch::TensorOptions().dtype(torch::kInt64).device(torch::kCUDA, device);
    auto o32 = torch::TensorOptions().dtype(torch::kInt32).device(torch::kCUDA, device);
    auto o8 = torch::TensorOptions().dtype(torch::kUInt8).device(torch::kCUDA, device);
    counters = torch::zeros({n_types * 8 + n_rel * 4 + 8}, o32);   // [types | relations | overflow, scratch]
    step = torch::zeros({1}, o32);
    int32_t* cbase = counters.data_ptr<int32_t>();
    // id tables: key arrays carved out of one buffer so a batch starts with ONE memset
    int64_t total_slots = 0;
    std::vector<int64_t> slots(n_types);
    cap_nodes.assign(n_types, 0);
    int64_t max_seed_all = 1;
    for (int t = 0; t < n_types; ++t) {
      int64_t cn = 0;
      for (int h = 0; h <= hops; ++h) cn += cap_rows[t][h];
      if (num_nodes[t] > 0) cn = std::min(cn, num_nodes[t] + max_seeds[t]);
      cap_nodes[t] = std::max<int64_t>(cn, 16);
      TORCH_CHECK(cap_nodes[t] < (1LL << 30), "sampler arena too large");
      int64_t sl = 64;
      while (sl < cap_nodes[t] * 2) sl <<= 1;
      slots[t] = sl;
      total_slots += sl;
      max_seed_all = std::max(max_seed_all, max_seeds[t]);
    }
    all_keys = torch::full({total_slots}, -1, o64);
    int64_t off = 0;
    for (int t = 0; t < n_types; ++t) {
      auto tb = std::make_unique<DeviceTable>(device, 16);   // small placeholder, re-pointed below
      tb->cap_nodes = cap_nodes[t];
      tb->keys = all_keys.narrow(0, off, slots[t]);
      tb->vals = torch::zeros({slots[t]}, o32);
      tb->aux = torch::zeros({slots[t]}, o32);
      tb->nodes = torch::zeros({cap_nodes[t]}, o64);
      tb->cursor = counters.narrow(0, ctr_type_base(t) + 6, 1);
      tb->ht.keys = tb->keys.data_ptr<int64_t>();
      tb->ht.vals = tb->vals.data_ptr<int32_t>();
      tb->ht.aux = tb->aux.data_ptr<int32_t>();
      tb->ht.mask = static_cast<uint32_t>(slots[t] - 1);
      off += slots[t];
      tables.push_back(std::move(tb));
    }
    seed_scratch = torch::zeros({max_seed_all}, o32);
    seed_local = torch::zeros({n_types, max_seed_all}, o32);
    // per-relation degree arrays (indexed by the key type's local id), one contiguous buffer
    int64_t deg_total = 0;
    for (int r = 0; r < n_rel; ++r) deg_total += cap_nodes[key_type[r]];
    deg_all = torch::zeros({deg_total}, o32);
    off = 0;
    for (int r = 0; r < n_rel; ++r) {
      deg.push_back(deg_all.narrow(0, off, cap_nodes[key_type[r]]));
      off += cap_nodes[key_type[r]];
    }
    ell.resize(n_rel);
    ell_eids.resize(n_rel);
    for (int r = 0; r < n_rel; ++r)
      for (int h = 0; h < hops; ++h) {
        const int64_t n = cap_rows[key_type[r]][h] * fanouts[r][h];
        ell[r].push_back(torch::full({std::max<int64_t>(n, 1)}, -1, o32));
        ell_eids[r].push_back(with_edge ? torch::full({std::max<int64_t>(n, 1)}, -1, o64) : Tensor());
      }
    // static launch descriptors
    std::vector<HopArgs> hd(static_cast<size_t>(hops) * n_rel);
    max_k.assign(hops, 0);
    max_rows.assign(hops, 1);
    for (int h = 0; h < hops; ++h)
      for (int r = 0; r < n_rel; ++r) {
        HopArgs a{};
        const int kt = key_type[r], nt = nbr_type[r];
        a.g = graphs[r]->tbl;
        a.t = tables[nt]->ht;
        a.c.cum = cbase + ctr_type_base(kt);
        a.c.edges = cbase + ctr_rel_base(r);
        a.c.cursor = cbase + ctr_type_base(nt) + 6;
        a.c.overflow = cbase + ctr_overflow();
        a.nodes = tables[kt]->nodes.data_ptr<int64_t>();
        a.nodes_out = tables[nt]->nodes.data_ptr<int64_t>();
        a.ell = ell[r][h].data_ptr<int32_t>();
        a.ell_eids = with_edge ? ell_eids[r][h].data_ptr<int64_t>() : nullptr;
        a.deg = deg[r].data_ptr<int32_t>();
        a.hop = h;
        a.k = static_cast<int>(fanouts[r][h]);
        a.cap_rows = static_cast<int>(cap_rows[kt][h]);
        a.cap_nodes = static_cast<int>(cap_nodes[nt]);
        a.cap_rows_next = static_cast<int>(cap_rows[nt][h + 1]);
        a.weighted = weighted ? 1 : 0;
        a.replace = replace ? 1 : 0;
        a.seed = static_cast<uint64_t>(seed);
        a.stream = static_cast<uint32_t>(h * n_rel + r);
        a.stream_dev = step.data_ptr<int32_t>();
        a.stream_stride = 1;   // `step` holds the Philox stream base itself (callers add hops * n_rel per batch)
        a.bound_ptr = cbase + ctr_type_base(nt) + h + 2;
        hd[static_cast<size_t>(h) * n_rel + r] = a;
        if (a.k > 0) {
          max_k[h] = std::max(max_k[h], a.k);
          max_rows[h] = std::max(max_rows[h], a.cap_rows);
        }
      }
    descs = torch::empty({static_cast<int64_t>(hd.size() * sizeof(HopArgs))}, o8);
    cudaMemcpy(descs.data_ptr(), hd.data(), hd.size() * sizeof(HopArgs), cudaMemcpyHostToDevice);
    std::vector<HeteroTypeState> ts(n_types);
    for (int t = 0; t < n_types; ++t) {
      ts[t].cum = cbase + ctr_type_base(t);
      ts[t].cursor = cbase + ctr_type_base(t) + 6;
      ts[t].cap_nodes = static_cast<int>(cap_nodes[t]);
      for (int h = 0; h < 5; ++h) ts[t].cap_rows[h] = h <= hops ? static_cast<int>(cap_rows[t][h]) : 0;
    }
    type_states = torch::empty({static_cast<int64_t>(ts.size() * sizeof(HeteroTypeState))}, o8);
    cudaMemcpy(type_states.data_ptr(), ts.data(), ts.size() * sizeof(HeteroTypeState), cudaMemcpyHostToDevice);
    check_cuda_err("HeteroArena setup");
  }

  // seeds of one or more node types -> multi-hop sample of every relation; no host synchronisation
  void sample(const std::vector<int64_t>& seed_types, const std::vector<Tensor>& seeds, int64_t step_inc) {
    c10::cuda::CUDAGuard guard(device);
    TORCH_CHECK(seed_types.size() == seeds.size() && !seeds.empty());
    cudaStream_t s = cur_stream();
    int32_t* cbase = counters.data_ptr<int32_t>();
    cudaMemsetAsync(cbase, 0, sizeof(int32_t) * ctr_overflow(), s);          // every cum / cursor / edge counter
    cudaMemsetAsync(all_keys.data_ptr(), 0xFF, all_keys.numel() * sizeof(int64_t), s);
    if (any_zero_fanout) cudaMemsetAsync(deg_all.data_ptr(), 0, deg_all.numel() * sizeof(int32_t), s);
    for (size_t i = 0; i < seeds.size(); ++i) {
      const int t = static_cast<int>(seed_types[i]);
      const Tensor& sd = seeds[i];
      TORCH_CHECK(t >= 0 && t < n_types && sd.is_cuda() && sd.scalar_type() == torch::kInt64 && sd.is_contiguous());
      TORCH_CHECK(sd.numel() <= max_seeds[t], "more seeds than the arena was built for");
      BatchCounters c{cbase + ctr_type_base(t), cbase + ctr_overflow() + 1 /* scratch */, cbase + ctr_type_base(t) + 6,
                      cbase + ctr_overflow()};
      // init_seeds zeroes 4 "edge" counters: point them at scratch words that nothing else reads
      launch_init_seeds(sd.data_ptr<int64_t>(), sd.numel(), nullptr, tables[t]->ht, tables[t]->nodes.data_ptr<int64_t>(),
                        seed_local.data_ptr<int32_t>() + t * seed_local.size(1), seed_scratch.data_ptr<int32_t>(), c,
                        (i == 0 && step_inc != 0) ? step.data_ptr<int32_t>() : nullptr, static_cast<int>(step_inc), s);
    }
    const HopArgs* d = reinterpret_cast<const HopArgs*>(descs.data_ptr());
    const HeteroTypeState* ts = reinterpret_cast<const HeteroTypeState*>(type_states.data_ptr());
    for (int h = 0; h < hops; ++h) {
      launch_sample_hop_grouped(d + static_cast<size_t>(h) * n_rel, n_rel, max_k[h], max_rows[h], s);
      launch_hetero_finalize(ts, n_types, h, s);
      launch_relabel_hop_grouped(d + static_cast<size_t>(h) * n_rel, n_rel, max_k[h], max_rows[h], s);
    }
    check_cuda_err("hetero arena sample");
  }

  // PyG-shaped output (ONE host sync for the sizes):
  //   nodes[t] (global ids, hop-contiguous), per relation (rows = neighbour local ids, cols = key local ids, eids),
  //   num_sampled_nodes[t][hop], num_sampled_edges[rel][hop]
  std::tuple<std::vector<Tensor>, std::vector<Tensor>, std::vector<Tensor>, std::vector<Tensor>,
             std::vector<std::vector<int64_t>>, std::vector<std::vector<int64_t>>> to_coo() {
    c10::cuda::CUDAGuard guard(device);
    auto o64 = torch::TensorOptions().dtype(torch::kInt64).device(torch::kCUDA, device);
    Tensor host = counters.cpu();  // sync
    const int32_t* c = host.data_ptr<int32_t>();
    std::vector<Tensor> nodes_out, rows_out, cols_out, eids_out;
    std::vector<std::vector<int64_t>> nn(n_types), ne(n_rel);
    for (int t = 0; t < n_types; ++t) {
      const int32_t* cum = c + ctr_type_base(t);
      for (int h = 0; h <= hops; ++h) nn[t].push_back(cum[h + 1] - cum[h]);
      nodes_out.push_back(tables[t]->nodes.narrow(0, 0, cum[hops + 1]).clone());
    }
    for (int r = 0; r < n_rel; ++r) {
      const int32_t* cum = c + ctr_type_base(key_type[r]);
      int64_t E = 0;
      for (int h = 0; h < hops; ++h) { ne[r].push_back(c[ctr_rel_base(r) + h]); E += c[ctr_rel_base(r) + h]; }
      const int64_t T = cum[hops];
      Tensor rows = torch::empty({E}, o64), cols = torch::empty({E}, o64);
      Tensor eids = with_edge ? torch::empty({E}, o64) : Tensor();
      if (E > 0) {
        Tensor offs = torch::zeros({T + 1}, o64);
        offs.narrow(0, 1, T).copy_(deg[r].narrow(0, 0, T).to(torch::kInt64).cumsum(0));
        for (int h = 0; h < hops; ++h) {
          if (fanouts[r][h] <= 0) continue;
          launch_ell_to_coo(ell[r][h].data_ptr<int32_t>(), with_edge ? ell_eids[r][h].data_ptr<int64_t>() : nullptr,
                            deg[r].data_ptr<int32_t>(), offs.data_ptr<int64_t>(),
                            counters.data_ptr<int32_t>() + ctr_type_base(key_type[r]), h, fanouts[r][h],
                            cap_rows[key_type[r]][h], rows.data_ptr<int64_t>(), cols.data_ptr<int64_t>(),
                            with_edge ? eids.data_ptr<int64_t>() : nullptr, cur_stream());
        }
      }
      rows_out.push_back(rows); cols_out.push_back(cols); eids_out.push_back(eids);
    }
    check_cuda_err("hetero to_coo");
    return {nodes_out, rows_out, cols_out, eids_out, nn, ne};
  }

  int64_t overflow_index() const { return ctr_overflow(); }
  Tensor seed_local_of(int64_t t) { return seed_local[t]; }
  Tensor nodes_of(int64_t t) { return tables[t]->nodes; }
  Tensor deg_of(int64_t r) { return deg[r]; }
  Tensor ell_of(int64_t r, int64_t h) { return ell[r][h]; }
};

// ---------------------------------------------------------------------------
// RowTableHandle: multi-source row store (UnifiedTensor).
// ---------------------------------------------------------------------------
struct RowTableHandle {
  RowTable tbl{};
  std::vector<Tensor> keep;
  unsigned local_mask = 0;   // bit p: part p is memory of `device` itself (not a peer GPU's HBM / pinned host)
  int device;
  torch::ScalarType dtype = torch::kFloat32;
  int64_t width = 0;

  explicit RowTableHandle(int dev) : device(dev) { tbl.num_parts = 0; tbl.row_begin[0] = 0; }

  // remote = true: the rows live in a PEER GPU's HBM (IPC-mapped); cudaPointerGetAttributes cannot tell -- it reports
  // the importing device for memory opened with cudaIpcOpenMemHandle -- so the caller says so (parallel/partitioned.py)
  void append(const Tensor& part, bool remote = false) {
    TORCH_CHECK(tbl.num_parts < kMaxParts, "too many row-table parts");
    TORCH_CHECK(part.dim() >= 1 && part.is_contiguous(), "parts must be contiguous [rows, ...]");
    const int64_t w = part.numel() / std::max<int64_t>(part.size(0), 1);
    if (tbl.num_parts == 0) {
      dtype = part.scalar_type();
      width = part.dim() > 1 ? w : 1;
      tbl.row_bytes = width * part.element_size();
    }
    TORCH_CHECK(part.scalar_type() == dtype, "dtype mismatch between parts");
    TORCH_CHECK((part.dim() > 1 ? w : 1) == width || part.size(0) == 0, "row width mismatch");
    tbl.base[tbl.num_parts] = part.size(0) > 0 ? dev_ptr(part) : nullptr;
    if (part.size(0) > 0 && part.is_cuda() && !remote) {
      cudaPointerAttributes at{};
      if (cudaPointerGetAttributes(&at, part.data_ptr()) == cudaSuccess && at.type == cudaMemoryTypeDevice &&
          at.device == device)
        local_mask |= 1u << tbl.num_parts;
      else
        cudaGetLastError();
    }
    tbl.row_begin[tbl.num_parts + 1] = tbl.row_begin[tbl.num_parts] + part.size(0);
    tbl.num_parts++;
    keep.push_back(part);
  }

  int64_t num_rows() const { return tbl.row_begin[tbl.num_parts]; }

  Tensor gather(const Tensor& idx, const c10::optional<Tensor>& id2index, int64_t out_width) {
    c10::cuda::CUDAGuard guard(device);
    TORCH_CHECK(idx.is_cuda() && idx.scalar_type() == torch::kInt64 && idx.is_contiguous());
    if (out_width <= 0) out_width = width;
    TORCH_CHECK(out_width >= width);
    auto o = torch::TensorOptions().dtype(dtype).device(torch::kCUDA, device);
    Tensor out = out_width == width ? torch::empty({idx.numel(), width}, o)
                                    : torch::zeros({idx.numel(), out_width}, o);
    const int64_t* map = (id2index.has_value() && id2index->defined()) ? id2index->data_ptr<int64_t>() : nullptr;
    launch_gather_rows(tbl, idx.data_ptr<int64_t>(), map, idx.numel(), nullptr, out.data_ptr(),
                       out_width * out.element_size(), cur_stream(), map ? id2index->numel() : 0);
    check_cuda_err("gather_rows");
    return out;
  }

  // copy the rows of the batch's nodes that are NOT in this GPU's HBM into xcache[local id] (see launch_stage_remote_rows)
  void stage_remote_rows(const Tensor& nodes, const Tensor& counters, int64_t n_idx, Tensor xcache) {
    c10::cuda::CUDAGuard guard(device);
    TORCH_CHECK(nodes.is_cuda() && nodes.scalar_type() == torch::kInt64 && nodes.is_contiguous());
    TORCH_CHECK(xcache.is_cuda() && xcache.is_contiguous() &&
                xcache.numel() * xcache.element_size() >= nodes.numel() * tbl.row_bytes && tbl.row_bytes % 16 == 0);
    launch_stage_remote_rows(tbl, local_mask, nodes.data_ptr<int64_t>(), counters.data_ptr<int32_t>(), n_idx,
                             nodes.numel(), xcache.data_ptr(), cur_stream());
    check_cuda_err("stage_remote_rows");
  }
  bool all_local() const { return local_mask == ((tbl.num_parts >= 32) ? ~0u : ((1u << tbl.num_parts) - 1u)); }

  // rows are MXFP8 (d e4m3 bytes + d/32 UE8M0 scales + padding to 16 bytes): dequantised bf16 [n, d]
  Tensor gather_mxfp8(const Tensor& idx, int64_t d) {
    c10::cuda::CUDAGuard guard(device);
    TORCH_CHECK(dtype == torch::kUInt8 && d % 128 == 0 && width >= d + d / 32 && width % 16 == 0,
                "gather_mxfp8: table must hold uint8 rows of d + d/32 (+pad) bytes, d a multiple of 128");
    TORCH_CHECK(idx.is_cuda() && idx.scalar_type() == torch::kInt64 && idx.is_contiguous());
    Tensor out = torch::empty({idx.numel(), d}, torch::TensorOptions().dtype(torch::kBFloat16).device(torch::kCUDA, device));
    launch_gather_mxfp8(tbl, idx.data_ptr<int64_t>(), idx.numel(), d, out.data_ptr(), cur_stream());
    check_cuda_err("gather_mxfp8");
    return out;
  }

  // static-shape gather into a caller-provided buffer with a device-side count
  void gather_into(const Tensor& idx, const c10::optional<Tensor>& id2index,
                   const c10::optional<Tensor>& n_dev, Tensor out) {
    c10::cuda::CUDAGuard guard(device);
    const int64_t* map = (id2index.has_value() && id2index->defined()) ? id2index->data_ptr<int64_t>() : nullptr;
    const int32_t* nd = (n_dev.has_value() && n_dev->defined()) ? n_dev->data_ptr<int32_t>() : nullptr;
    TORCH_CHECK(out.dim() == 2 && out.stride(1) == 1 && out.size(0) >= idx.numel() && out.size(1) >= width &&
                (out.stride(0) * out.element_size()) % 16 == 0, "gather_into: [rows, >= width] destination with unit inner stride");
    launch_gather_rows(tbl, idx.data_ptr<int64_t>(), map, idx.numel(), nd, out.data_ptr(),
                       out.stride(0) * out.element_size(), cur_stream(), map ? id2index->numel() : 0);
    check_cuda_err("gather_into");
  }
};

// ---------------------------------------------------------------------------
// GraphSAGE engine kernels
// ---------------------------------------------------------------------------
static void fill_ell(const std::vector<Tensor>& ell, const std::vector<int64_t>& ks,
                     const int32_t** out_ell, int* out_k) {
  TORCH_CHECK(ell.size() == ks.size() && ell.size() <= 4);
  for (size_t i = 0; i < 4; ++i) { out_ell[i] = nullptr; out_k[i] = 1; }
  for (size_t i = 0; i < ell.size(); ++i) { out_ell[i] = ell[i].data_ptr<int32_t>(); out_k[i] = ks[i]; }
}

static SageAggArgs make_agg(RowTableHandle* feat, const c10::optional<Tensor>& nodes,
                            const c10::optional<Tensor>& src_local, int64_t d, const Tensor& counters,
                            int64_t n_hops_targets, int64_t cap_targets, const std::vector<Tensor>& ell,
                            const std::vector<int64_t>& ks, const Tensor& deg, void* out) {
  SageAggArgs a{};
  if (src_local.has_value() && src_local->defined()) {
    TORCH_CHECK(src_local->scalar_type() == torch::kBFloat16 && src_local->is_contiguous());
    a.src_local = src_local->data_ptr();
  } else {
    TORCH_CHECK(feat != nullptr && nodes.has_value(), "layer-1 aggregation needs a feature table + nodes");
    TORCH_CHECK(feat->dtype == torch::kBFloat16 || feat->dtype == torch::kUInt8,
                "engine features must be bf16 rows or MXFP8 (uint8) rows");
    TORCH_CHECK(feat->dtype == torch::kBFloat16 ? feat->width == d : feat->width == d + 16,
                "feature width mismatch (MXFP8 rows are d + 16 bytes)");
    a.feat = feat->tbl;
    a.nodes = nodes->data_ptr<int64_t>();
  }
  TORCH_CHECK(d % 8 == 0, "feature width must be a multiple of 8");
  a.d = d;
  a.cum = counters.data_ptr<int32_t>();
  a.n_hops_targets = n_hops_targets;
  a.cap_targets = cap_targets;
  fill_ell(ell, ks, a.ell, a.k);
  a.deg = deg.data_ptr<int32_t>();
  a.out = out;
  return a;
}

static void sage_aggregate(RowTableHandle* feat, const c10::optional<Tensor>& nodes,
                           const c10::optional<Tensor>& src_local, int64_t d, const Tensor& counters,
                           int64_t n_hops_targets, const std::vector<Tensor>& ell,
                           const std::vector<int64_t>& ks, const Tensor& deg, Tensor out) {
  c10::cuda::CUDAGuard guard(out.device());
  TORCH_CHECK(out.scalar_type() == torch::kBFloat16 && out.is_contiguous() && out.size(1) == 2 * d);
  TORCH_CHECK(feat == nullptr || (src_local.has_value() && src_local->defined()) || feat->dtype == torch::kBFloat16,
              "the unfused aggregation kernel reads bf16 rows; MXFP8 tables go through sage_fused");
  SageAggArgs a = make_agg(feat, nodes, src_local, d, counters, n_hops_targets, out.size(0), ell, ks,
                           deg, out.data_ptr());
  launch_sage_aggregate(a, cur_stream());
  check_cuda_err("sage_aggregate");
}

static void sage_scatter_bwd(const Tensor& dA, int64_t d, const Tensor& counters, int64_t n_hops_targets,
                             const std::vector<Tensor>& ell, const std::vector<int64_t>& ks,
                             const Tensor& deg, Tensor dH) {
  c10::cuda::CUDAGuard guard(dA.device());
  TORCH_CHECK(dA.scalar_type() == torch::kBFloat16 && dA.is_contiguous() && dA.size(1) == 2 * d);
  TORCH_CHECK(dH.scalar_type() == torch::kFloat32 && dH.is_contiguous() && dH.size(1) == d);
  SageScatterArgs a{};
  a.dA = dA.data_ptr();
  a.d = d;
  a.cum = counters.data_ptr<int32_t>();
  a.n_hops_targets = n_hops_targets;
  a.cap_targets = dA.size(0);
  fill_ell(ell, ks, a.ell, a.k);
  a.deg = deg.data_ptr<int32_t>();
  a.dH = dH.data_ptr<float>();
  launch_sage_scatter_bwd(a, cur_stream());
  check_cuda_err("sage_scatter_bwd");
}


// ---- column-block variants used by the heterogeneous engine: A_t = [mean_rel1 | mean_rel2 | ... | self] ----
static void sage_aggregate_block(RowTableHandle* feat, const c10::optional<Tensor>& nodes,
                                 const c10::optional<Tensor>& src_local, int64_t d, const Tensor& counters,
                                 int64_t n_hops_targets, const std::vector<Tensor>& ell,
                                 const std::vector<int64_t>& ks, const Tensor& deg, Tensor out, int64_t mean_col) {
  c10::cuda::CUDAGuard guard(out.device());
  TORCH_CHECK(out.scalar_type() == torch::kBFloat16 && out.dim() == 2 && out.stride(1) == 1 &&
              mean_col >= 0 && mean_col % 8 == 0 && mean_col + d <= out.size(1) && out.stride(0) % 8 == 0);
  SageAggArgs a = make_agg(feat, nodes, src_local, d, counters, n_hops_targets, out.size(0), ell, ks,
                           deg, out.data_ptr());
  a.out_ld = static_cast<int>(out.stride(0));
  a.mean_col = static_cast<int>(mean_col);
  a.self_col = -1;
  launch_sage_aggregate(a, cur_stream());
  check_cuda_err("sage_aggregate_block");
}

static void sage_scatter_block(const Tensor& dA, int64_t d, int64_t mean_col, const Tensor& counters,
                               int64_t n_hops_targets, const std::vector<Tensor>& ell,
                               const std::vector<int64_t>& ks, const Tensor& deg, Tensor dH) {
  c10::cuda::CUDAGuard guard(dA.device());
  TORCH_CHECK(dA.scalar_type() == torch::kBFloat16 && dA.dim() == 2 && dA.stride(1) == 1 && dA.stride(0) % 8 == 0 &&
              mean_col >= 0 && mean_col % 8 == 0 && mean_col + d <= dA.size(1));
  TORCH_CHECK(dH.scalar_type() == torch::kFloat32 && dH.is_contiguous() && dH.size(1) == d);
  SageScatterArgs a{};
  a.dA = dA.data_ptr();
  a.d = d;
  a.cum = counters.data_ptr<int32_t>();
  a.n_hops_targets = n_hops_targets;
  a.cap_targets = dA.size(0);
  fill_ell(ell, ks, a.ell, a.k);
  a.deg = deg.data_ptr<int32_t>();
  a.dH = dH.data_ptr<float>();
  a.dA_ld = static_cast<int>(dA.stride(0));
  a.mean_col = static_cast<int>(mean_col);
  a.self_col = -1;
  launch_sage_scatter_bwd(a, cur_stream());
  check_cuda_err("sage_scatter_block");
}

static void add_block_f32(const Tensor& dA, int64_t col, int64_t d, const Tensor& counters, int64_t n_hops, Tensor dH) {
  c10::cuda::CUDAGuard guard(dA.device());
  TORCH_CHECK(dA.scalar_type() == torch::kBFloat16 && dA.dim() == 2 && dA.stride(1) == 1 && dA.stride(0) % 8 == 0 &&
              col % 8 == 0 && col + d <= dA.size(1) && d % 8 == 0);
  TORCH_CHECK(dH.scalar_type() == torch::kFloat32 && dH.is_contiguous() && dH.size(1) == d);
  launch_add_block_f32(dA.data_ptr(), dA.stride(0), col, d, counters.data_ptr<int32_t>(), n_hops,
                       std::min<int64_t>(dA.size(0), dH.size(0)), dH.data_ptr<float>(), cur_stream());
  check_cuda_err("add_block_f32");
}

// EXPERIMENTAL atomics-free backward of the mean aggregation over the arena's transposed adjacency
static void sage_gather_bwd(const Tensor& dA, int64_t d, SamplerArena& ar, int64_t n_hops_targets,
                            const c10::optional<Tensor>& Z, Tensor dPre, const c10::optional<Tensor>& colsum,
                            bool prezeroed, double gscale) {
  c10::cuda::CUDAGuard guard(dA.device());
  TORCH_CHECK(ar.tr_hops >= n_hops_targets && n_hops_targets >= 1, "enable_transpose(n_hops) covers too few hops");
  TORCH_CHECK(dA.scalar_type() == torch::kBFloat16 && dA.is_contiguous() && dA.size(1) == 2 * d && d % 8 == 0);
  TORCH_CHECK(dPre.scalar_type() == torch::kBFloat16 && dPre.is_contiguous() && dPre.size(1) == d);
  TORCH_CHECK(d <= 1024, "sage_gather_bwd: d <= 1024");
  SageGatherBwdArgs a{};
  a.dA = dA.data_ptr();
  a.d = static_cast<int>(d);
  a.cum = ar.counters.data_ptr<int32_t>();
  a.n_hops_targets = static_cast<int>(n_hops_targets);
  a.cap_targets = static_cast<int>(dA.size(0));
  a.cap_src = static_cast<int>(dPre.size(0));
  TORCH_CHECK(a.cap_src <= ar.cap_nodes, "dPre has more rows than the arena has nodes");
  a.deg = ar.deg.data_ptr<int32_t>();
  a.off = ar.tr_off.data_ptr<int32_t>();
  a.cnt_upto = ar.tr_cnt.data_ptr<int32_t>() + (n_hops_targets - 1) * ar.cap_nodes;
  a.tgt = ar.tr_tgt.data_ptr<int32_t>();
  a.meta = ar.tr_meta.data_ptr<int32_t>();
  a.meta_inv = ar.tr_meta_inv.data_ptr<float>();
  a.Z = nullptr;
  if (Z.has_value() && Z->defined()) {
    TORCH_CHECK(Z->scalar_type() == torch::kBFloat16 && Z->is_contiguous() && Z->size(1) == d && Z->size(0) >= dPre.size(0));
    a.Z = Z->data_ptr();
  }
  a.dPre = dPre.data_ptr();
  a.colsum = nullptr;
  if (colsum.has_value() && colsum->defined()) {
    TORCH_CHECK(colsum->scalar_type() == torch::kFloat32 && colsum->numel() >= d && d % 4 == 0 &&
                reinterpret_cast<uintptr_t>(colsum->data_ptr()) % 16 == 0, "colsum: fp32 [d], 16-byte aligned");
    a.colsum = colsum->data_ptr<float>();
  }
  a.colsum_prezeroed = prezeroed ? 1 : 0;
  a.gscale = static_cast<float>(gscale);
  launch_sage_gather_bwd(a, cur_stream());
  check_cuda_err("sage_gather_bwd");
}

static void relu_bwd_cast(const Tensor& dH, const Tensor& Z, const Tensor& counters, int64_t n_hops,
                          Tensor dPre, const c10::optional<Tensor>& colsum, bool prezeroed, double gscale) {
  c10::cuda::CUDAGuard guard(dH.device());
  TORCH_CHECK(dPre.size(0) <= dH.size(0) && dPre.size(0) <= Z.size(0) && dPre.size(1) % 8 == 0);
  TORCH_CHECK(dH.scalar_type() == torch::kFloat32 && Z.scalar_type() == torch::kBFloat16 &&
              dPre.scalar_type() == torch::kBFloat16, "relu_bwd_cast: dH fp32, Z/dPre bf16");
  TORCH_CHECK(dH.is_contiguous() && Z.is_contiguous() && dPre.is_contiguous() && dH.size(1) == dPre.size(1) &&
              Z.size(1) == dPre.size(1), "relu_bwd_cast: contiguous [rows, d] operands of equal width");
  launch_relu_bwd_cast(dH.data_ptr<float>(), Z.data_ptr(), counters.data_ptr<int32_t>(), n_hops,
                       dPre.size(0), dPre.size(1), dPre.data_ptr(),
                       (colsum.has_value() && colsum->defined()) ? colsum->data_ptr<float>() : nullptr, cur_stream(),
                       prezeroed, static_cast<float>(gscale));
  check_cuda_err("relu_bwd_cast");
}

static void dropout_bf16(Tensor Z, const Tensor& counters, int64_t n_hops, double p, int64_t seed, int64_t layer,
                         const c10::optional<Tensor>& step_dev) {
  c10::cuda::CUDAGuard guard(Z.device());
  TORCH_CHECK(Z.scalar_type() == torch::kBFloat16 && Z.is_contiguous() && Z.dim() == 2 && Z.size(1) % 8 == 0,
              "dropout_bf16: contiguous bf16 [rows, d], d % 8 == 0");
  TORCH_CHECK(p >= 0.0 && p < 1.0, "dropout_bf16: p in [0, 1)");
  const int32_t* st = nullptr;
  if (step_dev.has_value() && step_dev->defined()) {
    TORCH_CHECK(step_dev->scalar_type() == torch::kInt32 && step_dev->is_cuda());
    st = step_dev->data_ptr<int32_t>();
  }
  launch_dropout_bf16(Z.data_ptr(), counters.data_ptr<int32_t>(), n_hops, Z.size(0), Z.size(1),
                      static_cast<float>(p), static_cast<uint64_t>(seed), static_cast<int>(layer), st, cur_stream());
  check_cuda_err("dropout_bf16");
}

static void bias_relu(Tensor Z, const Tensor& bias, const Tensor& counters, int64_t n_hops, bool relu) {
  c10::cuda::CUDAGuard guard(Z.device());
  TORCH_CHECK(Z.scalar_type() == torch::kBFloat16 && bias.scalar_type() == torch::kBFloat16);
  TORCH_CHECK(Z.size(1) % 8 == 0 && bias.numel() == Z.size(1));
  launch_bias_relu(Z.data_ptr(), bias.data_ptr(), counters.data_ptr<int32_t>(), n_hops, Z.size(0),
                   Z.size(1), relu, cur_stream());
  check_cuda_err("bias_relu");
}

static void softmax_nll(const Tensor& logits, int64_t C, const c10::optional<Tensor>& y,
                        const c10::optional<Tensor>& labels_all, const c10::optional<Tensor>& nodes,
                        const Tensor& counters, Tensor loss, Tensor dlogits,
                        const c10::optional<Tensor>& correct, const c10::optional<Tensor>& colsum, bool prezeroed) {
  c10::cuda::CUDAGuard guard(logits.device());
  TORCH_CHECK(logits.scalar_type() == torch::kBFloat16 && logits.is_contiguous());
  TORCH_CHECK(dlogits.sizes() == logits.sizes() && dlogits.is_contiguous());
  const bool direct = y.has_value() && y->defined();
  TORCH_CHECK(direct || (labels_all.has_value() && nodes.has_value()), "need y or (labels_all, nodes)");
  launch_softmax_nll(logits.data_ptr(), logits.size(1), C, direct ? y->data_ptr<int64_t>() : nullptr,
                     direct ? nullptr : labels_all->data_ptr<int64_t>(),
                     direct ? nullptr : nodes->data_ptr<int64_t>(),
                     counters.data_ptr<int32_t>(), logits.size(0), loss.data_ptr<float>(),
                     dlogits.data_ptr(),
                     (correct.has_value() && correct->defined()) ? correct->data_ptr<int32_t>() : nullptr,
                     (colsum.has_value() && colsum->defined()) ? colsum->data_ptr<float>() : nullptr,
                     cur_stream(), prezeroed);
  check_cuda_err("softmax_nll");
}

static void zero_grads(Tensor g, const c10::optional<Tensor>& loss, const c10::optional<Tensor>& correct) {
  c10::cuda::CUDAGuard guard(g.device());
  TORCH_CHECK(g.scalar_type() == torch::kFloat32 && g.is_contiguous() &&
              reinterpret_cast<uintptr_t>(g.data_ptr()) % 16 == 0);
  launch_zero_grads(g.data_ptr<float>(), g.numel(),
                    (loss.has_value() && loss->defined()) ? loss->data_ptr<float>() : nullptr,
                    (correct.has_value() && correct->defined()) ? correct->data_ptr<int32_t>() : nullptr, cur_stream());
  check_cuda_err("zero_grads");
}

static void colsum_bf16(const Tensor& X, const Tensor& counters, int64_t n_hops, Tensor out) {
  c10::cuda::CUDAGuard guard(X.device());
  TORCH_CHECK(X.scalar_type() == torch::kBFloat16 && X.is_contiguous() && X.size(1) % 8 == 0);
  TORCH_CHECK(out.scalar_type() == torch::kFloat32 && out.numel() == X.size(1) && X.size(1) <= 2048);
  launch_colsum_bf16(X.data_ptr(), counters.data_ptr<int32_t>(), n_hops, X.size(0), X.size(1),
                     out.data_ptr<float>(), cur_stream());
  check_cuda_err("colsum_bf16");
}

static void zero_rows(Tensor p, const Tensor& counters, int64_t n_hops) {
  c10::cuda::CUDAGuard guard(p.device());
  TORCH_CHECK(p.scalar_type() == torch::kFloat32 && p.is_contiguous() && p.size(1) % 4 == 0);
  launch_zero_rows(p.data_ptr<float>(), counters.data_ptr<int32_t>(), n_hops, p.size(0), p.size(1), cur_stream());
  check_cuda_err("zero_rows");
}

static void adam_step(Tensor p, const Tensor& g, Tensor m, Tensor v, const c10::optional<Tensor>& p_bf16,
                      double lr, double b1, double b2, double eps, double wd, const Tensor& step_dev,
                      double gscale) {
  c10::cuda::CUDAGuard guard(p.device());
  TORCH_CHECK(step_dev.scalar_type() == torch::kInt32 && step_dev.numel() >= 2,
              "step_dev must be int32[2]: {steps taken, block ticket}");
  TORCH_CHECK(p.scalar_type() == torch::kFloat32 && g.scalar_type() == torch::kFloat32 &&
              m.scalar_type() == torch::kFloat32 && v.scalar_type() == torch::kFloat32, "adam: fp32 state");
  TORCH_CHECK(g.numel() >= p.numel() && m.numel() >= p.numel() && v.numel() >= p.numel() && p.is_contiguous() &&
              g.is_contiguous() && m.is_contiguous() && v.is_contiguous(), "adam: state size / layout mismatch");
  launch_adam(p.data_ptr<float>(), g.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(),
              (p_bf16.has_value() && p_bf16->defined()) ? p_bf16->data_ptr() : nullptr, p.numel(), lr,
              b1, b2, eps, wd, step_dev.data_ptr<int32_t>(), gscale, cur_stream());
  check_cuda_err("adam");
}

static void sage_fused(RowTableHandle* feat, const c10::optional<Tensor>& nodes,
                       const c10::optional<Tensor>& src_local, int64_t d, const Tensor& counters,
                       int64_t n_hops_targets, const std::vector<Tensor>& ell,
                       const std::vector<int64_t>& ks, const Tensor& deg, const Tensor& w_packed,
                       const Tensor& bias, bool relu, Tensor z, const c10::optional<Tensor>& a_save,
                       const c10::optional<Tensor>& xcache) {
  c10::cuda::CUDAGuard guard(z.device());
  const bool fp8 = feat != nullptr && !(src_local.has_value() && src_local->defined()) && feat->dtype == torch::kUInt8;
  TORCH_CHECK(!fp8 || d == 128, "the MXFP8 loader of the fused kernel supports d = 128");
  TORCH_CHECK(z.scalar_type() == torch::kBFloat16 && z.is_contiguous());
  const int64_t n_out = z.size(1);
  TORCH_CHECK(sage_fused_supported(d, n_out), "unsupported fused shape d=", d, " n_out=", n_out);
  TORCH_CHECK(w_packed.numel() == 2 * d * n_out && w_packed.scalar_type() == torch::kBFloat16);
  SageFusedArgs f{};
  f.agg = make_agg(feat, nodes, src_local, d, counters, n_hops_targets, z.size(0), ell, ks, deg, nullptr);
  f.w_packed = w_packed.data_ptr();
  f.bias = bias.data_ptr();
  f.n_out = n_out;
  f.relu = relu;
  f.z = z.data_ptr();
  f.feat_fp8 = fp8 ? 1 : 0;
  // Measured on B200 (profiles/fused_trace_r2_{prefetch,noprefetch}.txt): the resolver-side prefetch.global.L2 of the next
  // tile's rows does NOT shorten the loaders' tile time (8.9 -> 9.3 us) and lengthens the prologue (4.2 -> 6.8 us): the
  // loaders are bound by bytes in flight, not by DRAM-vs-L2 latency.  Off by default; GLT_B200_L2_PREFETCH=1 enables it.
  static const bool l2pf = [] { const char* e = std::getenv("GLT_B200_L2_PREFETCH"); return e ? std::atoi(e) != 0 : false; }();
  f.l2_prefetch = (l2pf && feat != nullptr && !(src_local.has_value() && src_local->defined())) ? 1 : 0;
  f.local_mask = feat != nullptr ? feat->local_mask : 0u;
  f.xcache = nullptr;
  f.cache_part = -1;
  if (feat != nullptr && xcache.has_value() && xcache->defined()) {
    // remote rows of the batch were staged into `xcache` (indexed by local node id): reuse a remote part's slot
    for (int p = 0; p < feat->tbl.num_parts; ++p)
      if (!((feat->local_mask >> p) & 1u)) { f.cache_part = p; break; }
    if (f.cache_part >= 0) {
      TORCH_CHECK(xcache->is_cuda() && xcache->is_contiguous() &&
                  xcache->numel() * xcache->element_size() >= static_cast<int64_t>(nodes->numel()) * feat->tbl.row_bytes,
                  "sage_fused: xcache must hold one feature row per arena node");
      f.xcache = xcache->data_ptr();
    }
  }
  f.a_save = nullptr;
  if (a_save.has_value() && a_save->defined()) {
    TORCH_CHECK(a_save->size(0) >= z.size(0) && a_save->size(1) == 2 * d && a_save->is_contiguous());
    f.a_save = a_save->data_ptr();
  }
  launch_sage_fused(f, at::cuda::getCurrentDeviceProperties()->multiProcessorCount, cur_stream());
  check_cuda_err("sage_fused");
}


// ---------------------------------------------------------------------------
// TcGemm: a cached launch of the TMA-fed tcgen05 GEMM kernel (csrc/cuda/tc_gemm.cu) over fixed engine buffers.
// Tensor maps are encoded once; the batch-dependent extent is read from the sampler's device counters, so
// run() is a single static launch (CUDA-graph capturable).  Up to two problems share a launch.
// ---------------------------------------------------------------------------
struct TcGemm {
  TcGemmLaunch L;
  int device;
  std::vector<Tensor> keep;   // keeps every operand alive
  explicit TcGemm(int dev) : device(dev) {
    std::memset(&L, 0, sizeof(L));
  }
  static int pick_bn(int64_t n) { return n % 256 == 0 ? 256 : (n % 128 == 0 ? 128 : 64); }
  void check_bf16(const Tensor& t, const char* what) {
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == torch::kBFloat16 && t.dim() == 2 && t.stride(1) == 1 &&
                (t.stride(0) * 2) % 16 == 0 && reinterpret_cast<uintptr_t>(t.data_ptr()) % 16 == 0,
                "TcGemm: ", what, " must be a 2-D bf16 CUDA tensor with unit inner stride and 16-byte aligned rows");
  }
  void tmap(int slot, const Tensor& t, int box_cols, int box_rows) {
    const int rc = make_tmap_bf16_2d(L.maps[slot], t.data_ptr(), t.size(0), t.size(1), t.stride(0), box_cols, box_rows);
    TORCH_CHECK(rc == 0, "TcGemm: cuTensorMapEncodeTiled failed (rc=", rc, ")");
    keep.push_back(t);
  }
  TcProblem& next() {
    TORCH_CHECK(L.args.n_prob < 2, "TcGemm: at most two problems per launch");
    return L.args.p[L.args.n_prob];
  }
  void set_dyn(TcProblem& p, const Tensor& counters, int64_t idx, int64_t cap, bool is_k) {
    TORCH_CHECK(counters.is_cuda() && counters.scalar_type() == torch::kInt32 && idx >= 0 && idx < counters.numel());
    p.dyn = counters.data_ptr<int32_t>();
    p.dyn_idx = static_cast<int>(idx);
    p.dyn_cap = static_cast<int>(cap);
    p.dyn_is_k = is_k ? 1 : 0;
    keep.push_back(counters);
  }
  // Z[rows, N] = act(A[rows, K] . W[N, K]^T + bias); rows = min(counters[dyn_idx], A.size(0))
  void add_forward(const Tensor& A, const Tensor& W, const c10::optional<Tensor>& bias, bool relu, Tensor Z,
                   const Tensor& counters, int64_t dyn_idx) {
    c10::cuda::CUDAGuard guard(A.device());
    check_bf16(A, "A"); check_bf16(W, "W"); check_bf16(Z, "Z");
    const int64_t K = A.size(1), N = W.size(0);
    TORCH_CHECK(W.size(1) == K && Z.size(1) == N && Z.size(0) >= A.size(0) && N % 64 == 0 && K % 8 == 0);
    const int slot = 3 * L.args.n_prob;
    TcProblem& p = next();
    p.a_mn = 0; p.b_mn = 0; p.epi = 0; p.bn = pick_bn(N);
    p.m = A.size(0); p.n = N; p.k = K;
    set_dyn(p, counters, dyn_idx, A.size(0), false);
    p.bias = nullptr;
    if (bias.has_value() && bias->defined()) {
      TORCH_CHECK(bias->scalar_type() == torch::kBFloat16 && bias->numel() == N &&
                  reinterpret_cast<uintptr_t>(bias->data_ptr()) % 16 == 0);
      p.bias = bias->data_ptr();
      keep.push_back(*bias);
    }
    p.relu = relu ? 1 : 0;
    tmap(slot + 0, A, 64, 128);
    tmap(slot + 1, W, 64, p.bn);
    tmap(slot + 2, Z, 64, 128);
    L.max_items += static_cast<int>((A.size(0) + 127) / 128 * (N / p.bn));
    ++L.args.n_prob;
  }
  // dA[rows, N] = dPre[rows, Kd] . W[Kd, N]  (W row-major = N contiguous: the MN-major B operand)
  void add_dgrad(const Tensor& dPre, const Tensor& W, Tensor dA, const Tensor& counters, int64_t dyn_idx) {
    c10::cuda::CUDAGuard guard(dPre.device());
    check_bf16(dPre, "dPre"); check_bf16(W, "W"); check_bf16(dA, "dA");
    const int64_t Kd = dPre.size(1), N = W.size(1);
    TORCH_CHECK(W.size(0) == Kd && dA.size(1) == N && dA.size(0) >= dPre.size(0) && N % 64 == 0 && Kd % 8 == 0);
    const int slot = 3 * L.args.n_prob;
    TcProblem& p = next();
    p.a_mn = 0; p.b_mn = 1; p.epi = 0; p.bn = pick_bn(N);
    p.m = dPre.size(0); p.n = N; p.k = Kd;
    set_dyn(p, counters, dyn_idx, dPre.size(0), false);
    p.bias = nullptr; p.relu = 0;
    tmap(slot + 0, dPre, 64, 128);
    tmap(slot + 1, W, 64, 64);
    tmap(slot + 2, dA, 64, 128);
    L.max_items += static_cast<int>((dPre.size(0) + 127) / 128 * (N / p.bn));
    ++L.args.n_prob;
  }
  // gW[M, N] += dPre[rows, M]^T . A[rows, N], rows = min(counters[dyn_idx], dPre.size(0)); fp32 split-K red-add
  // (gW must be zeroed by the caller; dPre rows beyond the batch must be zero up to its capacity)
  void add_wgrad(const Tensor& dPre, const Tensor& A, Tensor gW, const Tensor& counters, int64_t dyn_idx) {
    c10::cuda::CUDAGuard guard(dPre.device());
    check_bf16(dPre, "dPre"); check_bf16(A, "A");
    const int64_t M = dPre.size(1), N = A.size(1);
    TORCH_CHECK(A.size(0) >= dPre.size(0) && N % 64 == 0 && M % 8 == 0);
    TORCH_CHECK(gW.is_cuda() && gW.scalar_type() == torch::kFloat32 && gW.dim() == 2 && gW.size(0) == M &&
                gW.size(1) == N && gW.stride(1) == 1 && gW.stride(0) % 4 == 0 &&
                reinterpret_cast<uintptr_t>(gW.data_ptr()) % 16 == 0, "TcGemm: gW must be fp32 [M, N], 16-byte aligned");
    const int slot = 3 * L.args.n_prob;
    TcProblem& p = next();
    p.a_mn = 1; p.b_mn = 1; p.epi = 1; p.bn = pick_bn(N);
    p.m = M; p.n = N; p.k = dPre.size(0);
    set_dyn(p, counters, dyn_idx, dPre.size(0), true);
    p.bias = nullptr; p.relu = 0;
    p.out32 = gW.data_ptr<float>(); p.ld32 = static_cast<int>(gW.stride(0)); p.m_valid = static_cast<int>(M);
    keep.push_back(gW);
    tmap(slot + 0, dPre, 64, 64);
    tmap(slot + 1, A, 64, 64);
    L.max_items += 1 << 20;   // split-K fills the grid
    ++L.args.n_prob;
  }
  void run() {
    TORCH_CHECK(L.args.n_prob >= 1, "TcGemm: no problem added");
    c10::cuda::CUDAGuard guard(static_cast<c10::DeviceIndex>(device));
    launch_tc_gemm(L, at::cuda::getCurrentDeviceProperties()->multiProcessorCount, cur_stream());
    check_cuda_err("tc_gemm");
  }
};


// ---------------------------------------------------------------------------
// TcGemmMx: block-scaled MXFP8 forward GEMM (csrc/cuda/tc_gemm_mx.cu).  A / W are e4m3 bytes [rows, K]; sfa / sfb are
// the UE8M0 scales packed into the tensor core's 512-byte block layout (data/quantize.py pack_mx_scale_blocks).
// ---------------------------------------------------------------------------
struct TcGemmMx {
  alignas(64) unsigned char maps[3][128];
  TcMxArgs a{};
  int device;
  std::vector<Tensor> keep;
  TcGemmMx(int dev, const Tensor& A, const Tensor& sfa, const Tensor& W, const Tensor& sfb,
           const c10::optional<Tensor>& bias, bool relu, Tensor Z, const c10::optional<Tensor>& counters,
           int64_t dyn_idx) : device(dev) {
    c10::cuda::CUDAGuard guard(static_cast<c10::DeviceIndex>(dev));
    auto ok8 = [](const Tensor& t) {
      return t.is_cuda() && t.scalar_type() == torch::kUInt8 && t.dim() == 2 && t.stride(1) == 1 && t.stride(0) % 16 == 0 &&
             reinterpret_cast<uintptr_t>(t.data_ptr()) % 16 == 0;
    };
    TORCH_CHECK(ok8(A) && ok8(W), "TcGemmMx: A / W must be uint8 (e4m3) [rows, K] CUDA tensors, 16-byte aligned rows");
    const int64_t M = A.size(0), K = A.size(1), N = W.size(0);
    TORCH_CHECK(W.size(1) == K && K % 128 == 0 && N % 128 == 0 && M >= 128, "TcGemmMx: K, N multiples of 128, M >= 128");
    TORCH_CHECK(Z.is_cuda() && Z.scalar_type() == torch::kBFloat16 && Z.dim() == 2 && Z.size(0) >= M && Z.size(1) == N &&
                Z.stride(1) == 1);
    const int64_t mb = (M + 127) / 128, kb = K / 128;
    TORCH_CHECK(sfa.is_cuda() && sfa.scalar_type() == torch::kUInt8 && sfa.is_contiguous() && sfa.numel() == mb * kb * 512,
                "TcGemmMx: sfa must hold ceil(M/128) * K/128 blocks of 512 bytes");
    TORCH_CHECK(sfb.is_cuda() && sfb.scalar_type() == torch::kUInt8 && sfb.is_contiguous() && sfb.numel() == (N / 128) * kb * 512,
                "TcGemmMx: sfb must hold N/128 * K/128 blocks of 512 bytes");
    a.m = M; a.n = N; a.k = K;
    a.bn = N % 256 == 0 ? 256 : 128;
    a.dyn = nullptr; a.dyn_idx = 0;
    if (counters.has_value() && counters->defined()) {
      TORCH_CHECK(counters->scalar_type() == torch::kInt32 && dyn_idx >= 0 && dyn_idx < counters->numel());
      a.dyn = counters->data_ptr<int32_t>();
      a.dyn_idx = static_cast<int>(dyn_idx);
      keep.push_back(*counters);
    }
    a.sfa = sfa.data_ptr(); a.sfb = sfb.data_ptr();
    a.bias = nullptr;
    if (bias.has_value() && bias->defined()) {
      TORCH_CHECK(bias->scalar_type() == torch::kBFloat16 && bias->numel() == N);
      a.bias = bias->data_ptr();
      keep.push_back(*bias);
    }
    a.relu = relu ? 1 : 0;
    TORCH_CHECK(make_tmap_u8_2d(maps[0], A.data_ptr(), M, K, A.stride(0), 128, 128) == 0, "tensor map A");
    TORCH_CHECK(make_tmap_u8_2d(maps[1], W.data_ptr(), N, K, W.stride(0), 128, a.bn) == 0, "tensor map W");
    TORCH_CHECK(make_tmap_bf16_2d(maps[2], Z.data_ptr(), Z.size(0), N, Z.stride(0), 64, 128) == 0, "tensor map Z");
    keep.insert(keep.end(), {A, sfa, W, sfb, Z});
  }
  void run() {
    c10::cuda::CUDAGuard guard(static_cast<c10::DeviceIndex>(device));
    launch_tc_gemm_mx(maps, a, at::cuda::getCurrentDeviceProperties()->multiProcessorCount, cur_stream());
    check_cuda_err("tc_gemm_mx");
  }
};

static void enable_peer_access(int dev, int peer) {
  if (dev == peer) return;
  c10::cuda::CUDAGuard guard(dev);
  int can = 0;
  cudaDeviceCanAccessPeer(&can, dev, peer);
  TORCH_CHECK(can, "GPU ", dev, " cannot access peer GPU ", peer);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return; }
  TORCH_CHECK(e == cudaSuccess, "cudaDeviceEnablePeerAccess: ", cudaGetErrorString(e));
}

// ---------------------------------------------------------------------------
// PeerBuffer: cudaMalloc'ed shard that other processes map with CUDA IPC *on their own
// device* (cudaIpcMemLazyEnablePeerAccess), so their kernels can dereference it over
// NVLink.  (Memory imported through torch's own IPC path is opened in the exporter
// device's context and is not reachable from kernels running on the importer's GPU.)
// ---------------------------------------------------------------------------
struct PeerBuffer : public std::enable_shared_from_this<PeerBuffer> {
  void* ptr = nullptr;
  size_t bytes = 0;
  int device = 0;       // device whose kernels may use `ptr`
  bool imported = false;

  ~PeerBuffer() {
    if (ptr == nullptr) return;
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(device);
    if (imported) cudaIpcCloseMemHandle(ptr); else cudaFree(ptr);
    cudaSetDevice(cur);
  }

  static std::shared_ptr<PeerBuffer> allocate(int device, int64_t nbytes) {
    c10::cuda::CUDAGuard guard(device);
    auto b = std::make_shared<PeerBuffer>();
    b->device = device;
    b->bytes = std::max<int64_t>(nbytes, 256);
    cudaError_t e = cudaMalloc(&b->ptr, b->bytes);
    TORCH_CHECK(e == cudaSuccess, "cudaMalloc(", b->bytes, ") failed: ", cudaGetErrorString(e));
    return b;
  }

  py::bytes handle() const {
    TORCH_CHECK(!imported, "only the owner can export a handle");
    c10::cuda::CUDAGuard guard(device);
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    TORCH_CHECK(e == cudaSuccess, "cudaIpcGetMemHandle failed: ", cudaGetErrorString(e));
    return py::bytes(reinterpret_cast<const char*>(&h), sizeof(h));
  }

  static std::shared_ptr<PeerBuffer> open(const std::string& handle_bytes, int my_device, int64_t nbytes) {
    TORCH_CHECK(handle_bytes.size() == sizeof(cudaIpcMemHandle_t), "bad IPC handle size");
    c10::cuda::CUDAGuard guard(my_device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle_bytes.data(), sizeof(h));
    auto b = std::make_shared<PeerBuffer>();
    b->device = my_device;
    b->bytes = nbytes;
    b->imported = true;
    cudaError_t e = cudaIpcOpenMemHandle(&b->ptr, h, cudaIpcMemLazyEnablePeerAccess);
    TORCH_CHECK(e == cudaSuccess, "cudaIpcOpenMemHandle failed: ", cudaGetErrorString(e));
    return b;
  }

  // non-owning tensor view (keeps the buffer alive through the deleter)
  Tensor as_tensor(torch::ScalarType dtype, std::vector<int64_t> sizes) {
    int64_t n = 1;
    for (auto v : sizes) n *= v;
    TORCH_CHECK(static_cast<size_t>(n) * c10::elementSize(dtype) <= bytes, "view exceeds the peer buffer");
    std::shared_ptr<PeerBuffer> keep = shared_from_this();
    auto opts = torch::TensorOptions().dtype(dtype).device(torch::kCUDA, device);
    return torch::from_blob(ptr, sizes, [keep](void*) mutable { keep.reset(); }, opts);
  }
};

// PeerGroup: the per-rank view of all ranks' gradient buffers + barrier flags.
struct PeerGroup {
  PeerPtrs p{};
  std::vector<Tensor> keep;
  Tensor epoch, err;
  int device;

  PeerGroup(int dev, int rank, std::vector<Tensor> grads, std::vector<Tensor> flags) : device(dev) {
    c10::cuda::CUDAGuard guard(dev);
    TORCH_CHECK(grads.size() == flags.size() && grads.size() <= kMaxParts);
    p.world = grads.size();
    p.rank = rank;
    for (size_t r = 0; r < grads.size(); ++r) {
      TORCH_CHECK(grads[r].scalar_type() == torch::kFloat32 && flags[r].scalar_type() == torch::kInt32);
      TORCH_CHECK(flags[r].numel() >= 2 * p.world);
      p.g[r] = grads[r].data_ptr<float>();
      p.flags[r] = flags[r].data_ptr<int32_t>();
      keep.push_back(grads[r]);
      keep.push_back(flags[r]);
    }
    auto o32 = torch::TensorOptions().dtype(torch::kInt32).device(torch::kCUDA, dev);
    epoch = torch::zeros({2}, o32);
    err = torch::zeros({1}, o32);
  }

  void barrier(int64_t which) {
    c10::cuda::CUDAGuard guard(device);
    launch_peer_barrier(p, which, epoch.data_ptr<int32_t>(), err.data_ptr<int32_t>(), cur_stream());
    check_cuda_err("peer_barrier");
  }

  void adam(Tensor param, Tensor m, Tensor v, const c10::optional<Tensor>& p_bf16, double lr, double b1, double b2,
            double eps, double wd, const Tensor& step_dev, double gscale) {
    c10::cuda::CUDAGuard guard(device);
    TORCH_CHECK(param.numel() % 4 == 0, "flat parameter buffer must be padded to a multiple of 4");
    TORCH_CHECK(step_dev.scalar_type() == torch::kInt32 && step_dev.numel() >= 2,
                "step_dev must be int32[2]: {steps taken, block ticket}");
    launch_adam_peer(p, param.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(),
                     (p_bf16.has_value() && p_bf16->defined()) ? p_bf16->data_ptr() : nullptr, param.numel(), lr,
                     b1, b2, eps, wd, step_dev.data_ptr<int32_t>(), gscale, cur_stream(), err.data_ptr<int32_t>());
    check_cuda_err("adam_peer");
  }
};

static void multimem_copy(const Tensor& src, int64_t mc_ptr, int64_t dst_byte_offset) {
  c10::cuda::CUDAGuard guard(src.device());
  TORCH_CHECK(src.is_cuda() && src.is_contiguous());
  const int64_t nbytes = src.numel() * src.element_size();
  TORCH_CHECK(nbytes % 16 == 0 && dst_byte_offset % 16 == 0 && mc_ptr != 0);
  launch_multimem_copy(src.data_ptr(), reinterpret_cast<void*>(mc_ptr + dst_byte_offset), nbytes, cur_stream());
  check_cuda_err("multimem_copy");
}

static Tensor pack_weight(const Tensor& w) {
  c10::cuda::CUDAGuard guard(w.device());
  TORCH_CHECK(w.scalar_type() == torch::kBFloat16 && w.is_contiguous() && w.dim() == 2);
  TORCH_CHECK(w.size(1) % 64 == 0 && w.size(0) % 16 == 0, "W must be [N%16, K%64]");
  Tensor out = torch::empty_like(w);
  launch_pack_weight(w.data_ptr(), w.size(0), w.size(1), out.data_ptr(), cur_stream());
  check_cuda_err("pack_weight");
  return out;
}

static void pack_weight_into(const Tensor& w, Tensor out) {
  c10::cuda::CUDAGuard guard(w.device());
  TORCH_CHECK(w.scalar_type() == torch::kBFloat16 && w.is_contiguous() && w.dim() == 2);
  TORCH_CHECK(out.numel() == w.numel() && out.scalar_type() == torch::kBFloat16);
  launch_pack_weight(w.data_ptr(), w.size(0), w.size(1), out.data_ptr(), cur_stream());
  check_cuda_err("pack_weight");
}

}  // namespace glt

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  using namespace glt;
  m.doc() = "graphlearn_for_pytorch_b200 native core (CPU reference ops + sm_100a kernels)";

  // ---- CPU ops ----
  m.def("coo_to_csr", &coo_to_csr);
  m.def("cpu_sample_neighbors", &cpu_sample_neighbors);
  m.def("cpu_sample_neighbors_weighted", &cpu_sample_neighbors_weighted);
  m.def("cpu_negative_sample", &cpu_negative_sample);
  m.def("cpu_node_subgraph", &cpu_node_subgraph);
  m.def("cpu_random_walk", &cpu_random_walk);
  m.def("cpu_stitch", &cpu_stitch);
  m.def("cpu_nbr_prob", &cpu_nbr_prob);
  m.def("cpu_gather_rows", &cpu_gather_rows, py::arg("table"), py::arg("ids"), py::arg("id2index"), py::arg("offset"),
        py::arg("out"), py::arg("pos"));
  py::class_<CpuIdTable>(m, "CpuIdTable")
      .def(py::init<int64_t>())
      .def("reset", &CpuIdTable::reset)
      .def("insert", &CpuIdTable::insert)
      .def("lookup", &CpuIdTable::lookup)
      .def("keys", &CpuIdTable::keys, py::arg("start") = 0)
      .def("size", &CpuIdTable::size);

  // ---- shm sample queue ----
  py::register_exception<QueueTimeoutError>(m, "QueueTimeoutError");
  py::register_exception<QueueClosedError>(m, "QueueClosedError");
  py::class_<SampleQueue>(m, "SampleQueue")
      .def(py::init<size_t, size_t>(), py::arg("max_msgs"), py::arg("buf_bytes"))
      .def(py::init<const std::string&>(), py::arg("name"))
      .def_property_readonly("name", &SampleQueue::name)
      .def("send", &SampleQueue::send, py::call_guard<py::gil_scoped_release>())
      .def("recv", &SampleQueue::recv, py::arg("timeout_ms") = 0, py::call_guard<py::gil_scoped_release>())
      .def("empty", &SampleQueue::empty)
      .def("size", &SampleQueue::size)
      .def("close", &SampleQueue::close)
      .def("pin_memory", &SampleQueue::pin_memory);

  // ---- CUDA graph / tables ----
  py::class_<GraphHandle>(m, "GraphHandle")
      .def(py::init<int>())
      .def("add_shard", &GraphHandle::add_shard)
      .def_readonly("num_rows", &GraphHandle::num_rows)
      .def_readonly("has_eids", &GraphHandle::has_eids)
      .def_readonly("has_weights", &GraphHandle::has_weights)
      .def_property_readonly("num_parts", [](const GraphHandle& g) { return g.tbl.num_parts; })
      .def("sample_one_hop", &GraphHandle::sample_one_hop)
      .def("lookup_degree", &GraphHandle::lookup_degree)
      .def("full_neighbors", &GraphHandle::full_neighbors)
      .def("negative_sample", &GraphHandle::negative_sample)
      .def("random_walk", &GraphHandle::random_walk)
      .def("nbr_prob", &GraphHandle::nbr_prob);
  py::class_<DeviceTable>(m, "DeviceTable")
      .def(py::init<int, int64_t>())
      .def("clear", &DeviceTable::clear)
      .def("insert", &DeviceTable::insert)
      .def("init_ordered", &DeviceTable::init_ordered)
      .def("lookup", &DeviceTable::lookup)
      .def("size", &DeviceTable::size)
      .def("subgraph", &DeviceTable::subgraph)
      .def_readonly("nodes", &DeviceTable::nodes)
      .def_readonly("cursor", &DeviceTable::cursor)
      .def_readonly("capacity", &DeviceTable::cap_nodes);
  py::class_<SamplerArena>(m, "SamplerArena")
      .def(py::init<int, int64_t, std::vector<int64_t>, bool, int64_t, std::vector<int64_t>>(), py::arg("device"),
           py::arg("max_seeds"), py::arg("fanouts"), py::arg("with_edge"), py::arg("num_graph_nodes"),
           py::arg("cap_override") = std::vector<int64_t>{})
      .def("sample", &SamplerArena::sample, py::arg("graph"), py::arg("seeds"), py::arg("n_dev"), py::arg("seed"),
           py::arg("stream_base"), py::arg("weighted"), py::arg("replace"), py::arg("use_dev_step"),
           py::arg("step_inc") = 0)
      .def("to_coo", &SamplerArena::to_coo)
      .def_readonly("nodes", &SamplerArena::nodes)
      .def_readonly("deg", &SamplerArena::deg)
      .def_readonly("counters", &SamplerArena::counters)
      .def_readonly("seed_local", &SamplerArena::seed_local)
      .def_readonly("step", &SamplerArena::step)
      .def("enable_transpose", &SamplerArena::enable_transpose)
      .def_readwrite("deterministic", &SamplerArena::deterministic)
      .def_readonly("tr_hops", &SamplerArena::tr_hops)
      .def_readonly("tr_off", &SamplerArena::tr_off)
      .def_readonly("tr_cnt", &SamplerArena::tr_cnt)
      .def_readonly("tr_tgt", &SamplerArena::tr_tgt)
      .def_readonly("ell", &SamplerArena::ell)
      .def_readonly("ell_eids", &SamplerArena::ell_eids)
      .def_readonly("cap_rows", &SamplerArena::cap_rows)
      .def_readonly("cap_nodes", &SamplerArena::cap_nodes)
      .def_readonly("fanouts", &SamplerArena::fanouts);
  py::class_<HeteroArena>(m, "HeteroArena")
      .def(py::init<int, int64_t, std::vector<GraphHandle*>, std::vector<int64_t>, std::vector<int64_t>,
                    std::vector<std::vector<int64_t>>, std::vector<int64_t>, std::vector<int64_t>, bool, int64_t, bool,
                    bool, int64_t, std::vector<std::vector<int64_t>>>(),
           py::arg("device"), py::arg("n_types"), py::arg("graphs"), py::arg("key_type"), py::arg("nbr_type"),
           py::arg("fanouts"), py::arg("num_nodes"), py::arg("max_seeds"), py::arg("with_edge") = false,
           py::arg("seed") = 0, py::arg("weighted") = false, py::arg("replace") = false, py::arg("cap_limit") = 0,
           py::arg("cap_override") = std::vector<std::vector<int64_t>>{}, py::keep_alive<1, 4>())
      .def("sample", &HeteroArena::sample, py::arg("seed_types"), py::arg("seeds"), py::arg("step_inc") = 1)
      .def("to_coo", &HeteroArena::to_coo)
      .def("seed_local_of", &HeteroArena::seed_local_of)
      .def("overflow_index", &HeteroArena::overflow_index)
      .def("nodes_of", &HeteroArena::nodes_of)
      .def("deg_of", &HeteroArena::deg_of)
      .def("ell_of", &HeteroArena::ell_of)
      .def_readonly("counters", &HeteroArena::counters)
      .def_readonly("step", &HeteroArena::step)
      .def_readonly("cap_rows", &HeteroArena::cap_rows)
      .def_readonly("cap_nodes", &HeteroArena::cap_nodes)
      .def_readonly("hops", &HeteroArena::hops)
      .def_readonly("n_types", &HeteroArena::n_types)
      .def_readonly("n_rel", &HeteroArena::n_rel);
  py::class_<RowTableHandle>(m, "RowTableHandle")
      .def(py::init<int>())
      .def("append", &RowTableHandle::append, py::arg("part"), py::arg("remote") = false)
      .def("num_rows", &RowTableHandle::num_rows)
      .def_readonly("width", &RowTableHandle::width)
      .def_property_readonly("num_parts", [](const RowTableHandle& t) { return t.tbl.num_parts; })
      .def("gather", &RowTableHandle::gather, py::arg("idx"), py::arg("id2index") = py::none(),
           py::arg("out_width") = 0)
      .def("gather_into", &RowTableHandle::gather_into)
      .def("gather_mxfp8", &RowTableHandle::gather_mxfp8)
      .def("stage_remote_rows", &RowTableHandle::stage_remote_rows)
      .def("all_local", &RowTableHandle::all_local);

  // ---- GraphSAGE engine ----
  m.def("sage_aggregate", &sage_aggregate);
  m.def("sage_scatter_bwd", &sage_scatter_bwd);
  m.def("sage_aggregate_block", &sage_aggregate_block);
  m.def("sage_scatter_block", &sage_scatter_block);
  m.def("add_block_f32", &add_block_f32);
  m.def("relu_bwd_cast", &relu_bwd_cast, py::arg("dH"), py::arg("Z"), py::arg("counters"), py::arg("n_hops"),
        py::arg("dPre"), py::arg("colsum"), py::arg("prezeroed") = false, py::arg("gscale") = 1.0);
  m.def("dropout_bf16", &dropout_bf16, py::arg("Z"), py::arg("counters"), py::arg("n_hops"), py::arg("p"),
        py::arg("seed"), py::arg("layer"), py::arg("step_dev") = py::none());
  m.def("zero_grads", &zero_grads);
  m.def("set_pdl", [](bool on) { return set_pdl(on ? 1 : 0) != 0; });
  m.def("sage_gather_bwd", &sage_gather_bwd, py::arg("dA"), py::arg("d"), py::arg("arena"), py::arg("n_hops_targets"),
        py::arg("Z"), py::arg("dPre"), py::arg("colsum"), py::arg("prezeroed") = false, py::arg("gscale") = 1.0);
  m.def("bias_relu", &bias_relu);
  m.def("softmax_nll", &softmax_nll, py::arg("logits"), py::arg("C"), py::arg("y"), py::arg("labels_all"),
        py::arg("nodes"), py::arg("counters"), py::arg("loss"), py::arg("dlogits"), py::arg("correct"),
        py::arg("colsum"), py::arg("prezeroed") = false);
  m.def("adam_step", &adam_step);
  m.def("colsum_bf16", &colsum_bf16);
  m.def("zero_rows", &zero_rows);
  m.def("sage_fused", &sage_fused, py::arg("feat"), py::arg("nodes"), py::arg("src_local"), py::arg("d"),
        py::arg("counters"), py::arg("n_hops_targets"), py::arg("ell"), py::arg("ks"), py::arg("deg"),
        py::arg("w_packed"), py::arg("bias"), py::arg("relu"), py::arg("z"), py::arg("a_save") = py::none(),
        py::arg("xcache") = py::none());
  m.def("sage_fused_supported", &sage_fused_supported);
  m.def("sage_fused_trace", []() {
    // [148, 32] int64 clock64 timeline of the last fused launch made with GLT_B200_FUSED_TRACE=1
    Tensor t = torch::zeros({148, kFusedTraceSlots}, torch::kInt64);
    sage_fused_trace_copy(reinterpret_cast<unsigned long long*>(t.data_ptr<int64_t>()));
    return t;
  });
  py::class_<PeerBuffer, std::shared_ptr<PeerBuffer>>(m, "PeerBuffer")
      .def_static("allocate", &PeerBuffer::allocate)
      .def_static("open", &PeerBuffer::open)
      .def("handle", &PeerBuffer::handle)
      .def("as_tensor", &PeerBuffer::as_tensor)
      .def_readonly("bytes", &PeerBuffer::bytes)
      .def_readonly("device", &PeerBuffer::device)
      .def_readonly("imported", &PeerBuffer::imported);
  py::class_<PeerGroup>(m, "PeerGroup")
      .def(py::init<int, int, std::vector<Tensor>, std::vector<Tensor>>())
      .def("barrier", &PeerGroup::barrier)
      .def("adam", &PeerGroup::adam)
      .def_readonly("err", &PeerGroup::err)
      .def_readonly("epoch", &PeerGroup::epoch);
  py::class_<TcGemmMx>(m, "TcGemmMx")
      .def(py::init<int, const Tensor&, const Tensor&, const Tensor&, const Tensor&, const c10::optional<Tensor>&, bool,
                    Tensor, const c10::optional<Tensor>&, int64_t>(),
           py::arg("device"), py::arg("A"), py::arg("sfa"), py::arg("W"), py::arg("sfb"), py::arg("bias") = py::none(),
           py::arg("relu") = false, py::arg("Z"), py::arg("counters") = py::none(), py::arg("dyn_idx") = 0)
      .def("run", &TcGemmMx::run);
  py::class_<TcGemm>(m, "TcGemm")
      .def(py::init<int>())
      .def("add_forward", &TcGemm::add_forward)
      .def("add_dgrad", &TcGemm::add_dgrad)
      .def("add_wgrad", &TcGemm::add_wgrad)
      .def("run", &TcGemm::run);
  m.def("enable_peer_access", &enable_peer_access);
  m.def("multimem_copy", &multimem_copy);
  m.def("pack_weight", &pack_weight);
  m.def("pack_weight_into", &pack_weight_into);
}
