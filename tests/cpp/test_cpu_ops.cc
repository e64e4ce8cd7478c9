// Native unit tests of the CPU operators and the tensor-map serializer (libtorch, no Python interpreter, no GPU).
// Counterparts of the reference's test/cpp/{test_graph,test_random_sampler,test_inducer,test_subgraph,
// test_random_negative_sampler,test_stitch_sample_results,test_tensor_map_serializer}.cu -- same techniques:
// tiny hand-written graphs, set-membership assertions for random output, exact expectations for deterministic ops.
#include <torch/torch.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>

#include "cpu/cpu_ops.h"
#include "cpu/sample_queue.h"

using torch::Tensor;

#define CHECK_T(cond)                                                            \
  do {                                                                           \
    if (!(cond)) {                                                               \
      std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
      std::exit(1);                                                              \
    }                                                                            \
  } while (0)

static Tensor i64(std::vector<int64_t> v) { return torch::tensor(v, torch::kInt64); }
static std::vector<int64_t> vec(const Tensor& t) {
  Tensor c = t.contiguous().to(torch::kInt64);
  return std::vector<int64_t>(c.data_ptr<int64_t>(), c.data_ptr<int64_t>() + c.numel());
}

// 4 x 6 adjacency used by the reference's sampler tests: row r has r + 1 neighbours
struct Csr { Tensor indptr, indices, eids; };
static Csr tiny_graph() {
  std::vector<int64_t> rows, cols;
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c <= r; ++c) { rows.push_back(r); cols.push_back((r * 2 + c) % 6); }
  auto [ptr, ind, e, w] = glt::coo_to_csr(i64(rows), i64(cols), c10::nullopt, c10::nullopt, 4, true);
  (void)w;
  return {ptr, ind, e};
}

static void test_coo_to_csr() {
  // unsorted COO with explicit edge ids and weights: ids / weights must follow the permutation
  Tensor rows = i64({2, 0, 1, 0, 2}), cols = i64({1, 3, 0, 1, 0});
  Tensor eids = i64({10, 11, 12, 13, 14});
  Tensor w = torch::tensor({0.1f, 0.2f, 0.3f, 0.4f, 0.5f});
  auto [ptr, ind, oe, ow] = glt::coo_to_csr(rows, cols, eids, w, 3, true);
  CHECK_T(vec(ptr) == (std::vector<int64_t>{0, 2, 3, 5}));
  CHECK_T(vec(ind) == (std::vector<int64_t>{1, 3, 0, 0, 1}));
  CHECK_T(vec(oe) == (std::vector<int64_t>{13, 11, 12, 14, 10}));
  CHECK_T(std::abs(ow[0].item<float>() - 0.4f) < 1e-6 && std::abs(ow[4].item<float>() - 0.1f) < 1e-6);
  std::printf("coo_to_csr ok\n");
}

static void test_sampler() {
  Csr g = tiny_graph();
  Tensor seeds = i64({0, 1, 2, 3, 7});  // 7 is outside the CSR: no neighbours
  for (int k : {1, 2, 5}) {
    auto [nbrs, cnt, eids] = glt::cpu_sample_neighbors(g.indptr, g.indices, g.eids, seeds, k, true, false, 42, 0);
    auto c = vec(cnt);
    int64_t off = 0;
    for (int i = 0; i < 5; ++i) {
      const int64_t deg = i < 4 ? i + 1 : 0;
      CHECK_T(c[i] == std::min<int64_t>(deg, k));
      if (i < 4) {
        auto row = vec(g.indices.slice(0, g.indptr[i].item<int64_t>(), g.indptr[i + 1].item<int64_t>()));
        std::set<int64_t> truth(row.begin(), row.end()), seen;
        for (int64_t j = 0; j < c[i]; ++j) {
          const int64_t v = nbrs[off + j].item<int64_t>();
          CHECK_T(truth.count(v) == 1);                                   // sampled nbrs are true nbrs
          CHECK_T(seen.insert(v).second);                                  // without replacement
          const int64_t e = eids[off + j].item<int64_t>();
          CHECK_T(g.indices[(g.eids == e).nonzero()[0][0].item<int64_t>()].item<int64_t>() == v);  // edge id matches
        }
      }
      off += c[i];
    }
  }
  // determinism: the stream is a pure function of (seed, stream, row, draw)
  auto a = glt::cpu_sample_neighbors(g.indptr, g.indices, c10::nullopt, seeds, 2, false, false, 7, 3);
  auto b = glt::cpu_sample_neighbors(g.indptr, g.indices, c10::nullopt, seeds, 2, false, false, 7, 3);
  CHECK_T(torch::equal(std::get<0>(a), std::get<0>(b)));
  std::printf("cpu_sample_neighbors ok\n");
}

static void test_inducer() {
  glt::CpuIdTable t(8);
  CHECK_T(vec(t.insert(i64({5, 9, 5, 2}))) == (std::vector<int64_t>{0, 1, 0, 2}));      // first-seen order
  CHECK_T(vec(t.insert(i64({9, 7, -1, 2, 11}))) == (std::vector<int64_t>{1, 3, -1, 2, 4}));
  CHECK_T(vec(t.keys(0)) == (std::vector<int64_t>{5, 9, 2, 7, 11}) && vec(t.keys(3)) == (std::vector<int64_t>{7, 11}));
  CHECK_T(vec(t.lookup(i64({11, 100, 5}))) == (std::vector<int64_t>{4, -1, 0}));
  std::vector<int64_t> many(5000);
  for (int i = 0; i < 5000; ++i) many[i] = 1000003LL * i;                                 // forces rehashing
  t.insert(i64(many));
  CHECK_T(t.size() == 5005 && t.find_one(1000003LL * 4999) == 5004);
  t.reset();
  CHECK_T(t.size() == 0 && vec(t.insert(i64({3}))) == (std::vector<int64_t>{0}));
  // O(1) reset by generation stamp: keys of earlier generations are gone, recycled tables behave like new ones
  for (int round = 0; round < 50; ++round) {
    t.reset();
    CHECK_T(t.find_one(1000003LL * 7) == -1 && t.find_one(3) == -1);
    CHECK_T(vec(t.insert(i64({40 + round, 3, 40 + round}))) == (std::vector<int64_t>{0, 1, 0}));
    CHECK_T(vec(t.keys(0)) == (std::vector<int64_t>{40 + round, 3}));
  }
  std::printf("CpuIdTable (inducer) ok\n");
}

static void test_gather_rows_and_threads() {
  const int64_t n = 5000, w = 7;
  Tensor table = torch::arange(n * w, torch::kFloat32).view({n, w});
  Tensor ids = torch::randint(0, n, {20000}, torch::kInt64);
  Tensor out = torch::zeros({20000, w}, torch::kFloat32);
  at::set_num_threads(4);                                                          // chunks run on several threads
  glt::cpu_gather_rows(table, ids, c10::nullopt, 0, out, c10::nullopt);
  CHECK_T(torch::equal(out, table.index_select(0, ids)));
  Tensor perm = torch::randperm(n, torch::kInt64);                                  // id2index lookup + scatter positions
  Tensor pos = torch::randperm(20000, torch::kInt64);
  Tensor out2 = torch::zeros({20000, w}, torch::kFloat32);
  glt::cpu_gather_rows(table, ids, perm, 0, out2, pos);
  CHECK_T(torch::equal(out2.index_select(0, pos), table.index_select(0, perm.index_select(0, ids))));
  Tensor out3 = torch::zeros({3, w}, torch::kFloat32);                              // offset map (range partition book)
  glt::cpu_gather_rows(table, i64({1000, 1002, 1001}), c10::nullopt, 1000, out3, c10::nullopt);
  CHECK_T(torch::equal(out3, table.index_select(0, i64({0, 2, 1}))));
  bool raised = false;
  try { glt::cpu_gather_rows(table, i64({n}), c10::nullopt, 0, out3, c10::nullopt); } catch (const c10::Error&) { raised = true; }
  CHECK_T(raised);
  // the samplers give the same answer on one thread and on many (counter-based RNG, per-row output slots)
  auto [ptr, ind, e, wgt] = glt::coo_to_csr(torch::randint(0, 3000, {60000}, torch::kInt64),
                                            torch::randint(0, 3000, {60000}, torch::kInt64), c10::nullopt,
                                            c10::nullopt, 3000, true);
  (void)wgt;
  Tensor seeds = torch::randint(0, 3000, {40000}, torch::kInt64);
  at::set_num_threads(1);
  auto a = glt::cpu_sample_neighbors(ptr, ind, e, seeds, 5, true, false, 11, 3);
  at::set_num_threads(4);
  auto b = glt::cpu_sample_neighbors(ptr, ind, e, seeds, 5, true, false, 11, 3);
  CHECK_T(torch::equal(std::get<0>(a), std::get<0>(b)) && torch::equal(std::get<1>(a), std::get<1>(b)) &&
          torch::equal(std::get<2>(a), std::get<2>(b)));
  std::printf("cpu_gather_rows / threaded loops ok\n");
}

static void test_subgraph_and_negative() {
  Csr g = tiny_graph();
  auto [nodes, rows, cols, eids] = glt::cpu_node_subgraph(g.indptr, g.indices, g.eids, i64({3, 0, 2}), true);
  auto nv = vec(nodes), rv = vec(rows), cv = vec(cols);
  for (size_t i = 0; i < rv.size(); ++i) {     // every returned edge exists and both ends are in the node set
    const int64_t s = nv[rv[i]], d = nv[cv[i]];
    auto row = vec(g.indices.slice(0, g.indptr[s].item<int64_t>(), g.indptr[s + 1].item<int64_t>()));
    CHECK_T(std::find(row.begin(), row.end(), d) != row.end());
  }
  int64_t expect = 0;                           // and every induced edge is returned
  for (int64_t s : {3, 0, 2})
    for (int64_t d : vec(g.indices.slice(0, g.indptr[s].item<int64_t>(), g.indptr[s + 1].item<int64_t>())))
      expect += (d == 3 || d == 0 || d == 2);
  CHECK_T(static_cast<int64_t>(rv.size()) == expect);
  auto [nr, nc] = glt::cpu_negative_sample(g.indptr, g.indices, 4, 6, 50, 10, true, true, 1, 0);
  CHECK_T(nr.numel() == 50 && nc.numel() == 50);                        // padding: exactly req_num
  auto [sr, sc] = glt::cpu_negative_sample(g.indptr, g.indices, 4, 6, 20, 10, false, true, 1, 0);
  for (int64_t i = 0; i < sr.numel(); ++i) {   // strict: never an existing edge
    const int64_t s = sr[i].item<int64_t>(), d = sc[i].item<int64_t>();
    auto row = vec(g.indices.slice(0, g.indptr[s].item<int64_t>(), g.indptr[s + 1].item<int64_t>()));
    CHECK_T(std::find(row.begin(), row.end(), d) == row.end());
  }
  std::printf("cpu_node_subgraph / cpu_negative_sample ok\n");
}

static void test_stitch() {
  // two partitions answered for seeds {0,2} and {1}: exact merged arrays in seed order
  auto [nbrs, num, eids] = glt::cpu_stitch(3, {i64({0, 2}), i64({1})}, {i64({10, 11, 30}), i64({20, 21})},
                                           {i64({2, 1}), i64({2})}, {i64({100, 101, 300}), i64({200, 201})});
  CHECK_T(vec(num) == (std::vector<int64_t>{2, 2, 1}));
  CHECK_T(vec(nbrs) == (std::vector<int64_t>{10, 11, 20, 21, 30}));
  CHECK_T(vec(eids) == (std::vector<int64_t>{100, 101, 200, 201, 300}));
  std::printf("cpu_stitch ok\n");
}

static void test_serializer() {
  glt::SampleQueue q(4, 1 << 16);
  glt::TensorMap m;
  m["ids"] = i64({1, 2, 3});
  m["feat"] = torch::arange(12, torch::kFloat32).view({3, 4});
  m["half"] = torch::ones({2, 2}, torch::kFloat16);
  m["empty"] = torch::empty({0}, torch::kInt64);
  m["scalar"] = torch::tensor(7, torch::kInt32);
  q.send(m);
  CHECK_T(q.size() == 1);
  glt::TensorMap r = q.recv(1000);
  CHECK_T(r.size() == m.size());
  for (auto& kv : m) {
    CHECK_T(r.count(kv.first) == 1 && r[kv.first].scalar_type() == kv.second.scalar_type());
    CHECK_T(r[kv.first].sizes() == kv.second.sizes() && torch::equal(r[kv.first], kv.second));
  }
  bool timed_out = false;
  try { q.recv(50); } catch (const glt::QueueTimeoutError&) { timed_out = true; }
  CHECK_T(timed_out);
  std::printf("tensor-map serializer / SampleQueue ok\n");
}

int main() {
  test_coo_to_csr();
  test_sampler();
  test_inducer();
  test_gather_rows_and_threads();
  test_subgraph_and_negative();
  test_stitch();
  test_serializer();
  std::printf("CPU_OPS_OK\n");
  return 0;
}
