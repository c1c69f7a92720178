// shine_b200 — the compute-node process of the reference (src/main.cc, src/compute_node.cc) with the hot path
// replaced by libshn_b200.so (include/shn.h).  Kept contracts:
//   * CLI: every flag of IndexConfiguration (src/common/configuration.hh:57-113) and of the transport base class
//     (rdma-library/library/configuration.cc:17-93) is accepted with the same names, defaults and validation;
//     transport flags are parsed and ignored, `--servers` only fixes the number of dump parts (`_of<n>`).
//   * dataset directory: <data>/base.{fbin,u8bin,i8bin}, <data>/queries/{query,groundtruth,warmup}-<suffix>.*
//     (src/compute_node.cc:278-320, src/io/read_data.hh:21-36, src/io/deserializer.hh:24-45)
//   * index dumps: <data>/dump/index_m<M>_efc<efC>_node<i>_of<n>.dat (src/compute_node.cc:428-430)
//   * output: ONE JSON document on stdout with the reference's key tree (src/compute_node.cc:40-74,178-187,478-558;
//     src/common/statistics.hh:117-142), nlohmann-style dump(2) with sorted keys; status text on stderr
//   * errors: message on stderr + exit(EXIT_FAILURE) (rdma-library/library/utils.hh:17-23)
// `--is-server` (the memory-node role) has nothing to do here — the index lives in HBM — and exits 0.
#include <algorithm>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <thread>
#include <string>
#include <variant>
#include <vector>

#include "../../include/shn.h"

namespace fs = std::filesystem;

namespace {

[[noreturn]] void die(const std::string& msg) {
  std::cerr << "[ERROR]: " << msg << std::endl;
  std::exit(EXIT_FAILURE);
}
void status(const std::string& msg) { std::cerr << "[STATUS]: " << msg << std::endl; }  // compute_node.hh print_status

// ------------------------------------------------------------------------------------------------ configuration
struct Config {
  // transport (accepted, ignored)
  bool is_server = false, is_initiator = false;
  std::vector<std::string> servers, clients;
  uint32_t num_clients = 1, port = 1234, ib_port = 1;
  int max_poll_cqes = 16, max_send_wrs = 1024, max_recv_wrs = 1024;
  // index (configuration.hh:20-42)
  std::string data_path, query_suffix, label;
  uint32_t threads = 0, coroutines = 4;
  int seed = 1234;
  bool disable_thread_pinning = false, store_index = false, load_index = false, use_cache = false, routing = false,
       no_recall = false, ip_dist = false;
  uint32_t cache_ratio = 5, ef_search = 0, ef_construction = 200, k = 0, m = 32;
  int gpu = 0;   // extension: --gpu <ordinal>: first GPU to use
  int gpus = 1;  // extension: --gpus <n>: spread the index over n GPUs (memory nodes -> HBM partitions, NVLink peer reads)
};

[[noreturn]] void usage_exit(const char* argv0) {
  std::cerr << "Try " << argv0 << " --help" << std::endl;
  std::exit(EXIT_FAILURE);
}

Config parse(int argc, char** argv) {
  Config c;
  auto need = [&](int& i) -> std::string {
    if (i + 1 >= argc) { std::cerr << "[ERROR]: the required argument for option '" << argv[i] << "' is missing" << std::endl; usage_exit(argv[0]); }
    return argv[++i];
  };
  auto u32v = [&](int& i) { return static_cast<uint32_t>(std::stoul(need(i))); };
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i], inline_val;
    const size_t eq = a.find('=');
    bool has_inline = false;
    if (a.rfind("--", 0) == 0 && eq != std::string::npos) { inline_val = a.substr(eq + 1); a = a.substr(0, eq); has_inline = true; }
    auto val = [&]() { return has_inline ? inline_val : need(i); };
    auto multi = [&](std::vector<std::string>& out) {
      if (has_inline) out.push_back(inline_val);
      while (i + 1 < argc && argv[i + 1][0] != '-') out.push_back(argv[++i]);
    };
    try {
      if (a == "--help" || a == "-h") {
        std::cerr << "shine_b200: B200-native compute node; options as in the reference (src/common/configuration.hh, "
                     "rdma-library/library/configuration.cc) plus --gpu <ordinal> and --gpus <n>" << std::endl;
        std::exit(EXIT_FAILURE);
      } else if (a == "--is-server") c.is_server = true;
      else if (a == "--servers") multi(c.servers);
      else if (a == "--clients") multi(c.clients);
      else if (a == "--initiator" || a == "-i") c.is_initiator = true;
      else if (a == "--num-clients" || a == "-c") c.num_clients = static_cast<uint32_t>(std::stoul(val()));
      else if (a == "--port") c.port = static_cast<uint32_t>(std::stoul(val()));
      else if (a == "--ib-port") c.ib_port = static_cast<uint32_t>(std::stoul(val()));
      else if (a == "--max-poll-cqes") c.max_poll_cqes = std::stoi(val());
      else if (a == "--max-send-wrs") c.max_send_wrs = std::stoi(val());
      else if (a == "--max-receive-wrs") c.max_recv_wrs = std::stoi(val());
      else if (a == "--data-path" || a == "-d") c.data_path = val();
      else if (a == "--threads" || a == "-t") c.threads = static_cast<uint32_t>(std::stoul(val()));
      else if (a == "--coroutines" || a == "-C") c.coroutines = static_cast<uint32_t>(std::stoul(val()));
      else if (a == "--disable-thread-pinning" || a == "-p") c.disable_thread_pinning = true;
      else if (a == "--seed") c.seed = std::stoi(val());
      else if (a == "--label") c.label = val();
      else if (a == "--query-suffix" || a == "-q") c.query_suffix = val();
      else if (a == "--store-index" || a == "-s") c.store_index = true;  // `-s` is registered twice upstream; scripts use long forms
      else if (a == "--load-index" || a == "-l") c.load_index = true;
      else if (a == "--cache") c.use_cache = true;
      else if (a == "--routing") c.routing = true;
      else if (a == "--cache-ratio") c.cache_ratio = static_cast<uint32_t>(std::stoul(val()));
      else if (a == "--no-recall") c.no_recall = true;
      else if (a == "--ip-dist") c.ip_dist = true;
      else if (a == "--ef-search") c.ef_search = static_cast<uint32_t>(std::stoul(val()));
      else if (a == "--ef-construction") c.ef_construction = static_cast<uint32_t>(std::stoul(val()));
      else if (a == "-k" || a == "--k") c.k = static_cast<uint32_t>(std::stoul(val()));
      else if (a == "-m" || a == "--m") c.m = static_cast<uint32_t>(std::stoul(val()));
      else if (a == "--gpu") c.gpu = std::stoi(val());
      else if (a == "--gpus") c.gpus = std::stoi(val());
      else { std::cerr << "[ERROR]: unrecognised option '" << a << "'" << std::endl; usage_exit(argv[0]); }
    } catch (const std::exception& e) {
      std::cerr << "[ERROR]: the argument for option '" << a << "' is invalid" << std::endl;
      usage_exit(argv[0]);
    }
    (void)u32v;
  }
  // rdma-library/library/configuration.cc:63-83
  if (!c.is_server && c.servers.empty()) { std::cerr << "[ERROR]: --servers <arg-list> must be given if --is-server is not set" << std::endl; usage_exit(argv[0]); }
  if (c.is_server && c.is_initiator) { std::cerr << "[ERROR]: a server cannot be the initiator" << std::endl; usage_exit(argv[0]); }
  if (!c.is_initiator && !c.clients.empty()) { std::cerr << "[ERROR]: --clients <arg-list> is only required by the initiating client" << std::endl; usage_exit(argv[0]); }
  if (!c.is_server) {  // configuration.hh:88-113
    if (c.data_path.empty() || c.query_suffix.empty()) { std::cerr << "[ERROR]: Data path and query suffix cannot be empty" << std::endl; usage_exit(argv[0]); }
    if (c.threads == 0 || c.ef_search == 0 || c.k == 0) { std::cerr << "[ERROR]: Parameters threads, ef-search, and k are required" << std::endl; usage_exit(argv[0]); }
    if (c.store_index && c.load_index) { std::cerr << "[ERROR]: --store-index and --load-index cannot be used in conjunction" << std::endl; usage_exit(argv[0]); }
    if (c.use_cache && c.cache_ratio == 0) { std::cerr << "[ERROR]: If --cache is set, --cache-ratio must be > 0" << std::endl; usage_exit(argv[0]); }
    if (c.routing && !c.use_cache) { std::cerr << "[ERROR]: --routing can only be used in conjunction with --cache" << std::endl; usage_exit(argv[0]); }
  }
  return c;
}

// ------------------------------------------------------------------------------------------------ dataset files
struct Rows {
  uint32_t n = 0, dim = 0;
  std::vector<float> f;      // [n][dim] for vector files
  std::vector<uint32_t> u;   // [n][dim] for .bin (ground truth)
};

// `.fbin/.u8bin/.i8bin/.bin`: u32 n, u32 dim, n*dim elements row-major (io/read_data.hh:21-40, deserializer.hh:24-45)
Rows read_rows(const fs::path& file, bool meta_only = false) {
  std::ifstream in(file, std::ios::binary);
  if (!in) die("cannot open " + file.string());
  Rows r;
  in.read(reinterpret_cast<char*>(&r.n), 4);
  in.read(reinterpret_cast<char*>(&r.dim), 4);
  if (!in) die("Cannot read file " + file.string());
  const std::string ext = file.extension().string();
  if (ext != ".fbin" && ext != ".u8bin" && ext != ".i8bin" && ext != ".bin") die("unsupported file extension: " + ext);
  if (meta_only) return r;
  const size_t count = static_cast<size_t>(r.n) * r.dim;
  std::cerr << "reading input data... (dim=" << r.dim << ", num vectors=" << r.n << "/" << r.n << ", file=" << file.string() << ")" << std::endl;
  if (ext == ".fbin") {
    r.f.resize(count);
    in.read(reinterpret_cast<char*>(r.f.data()), count * 4);
  } else if (ext == ".bin") {
    r.u.resize(count);
    in.read(reinterpret_cast<char*>(r.u.data()), count * 4);
  } else {
    std::vector<uint8_t> raw(count);
    in.read(reinterpret_cast<char*>(raw.data()), count);
    r.f.resize(count);
    if (ext == ".u8bin") for (size_t i = 0; i < count; ++i) r.f[i] = static_cast<float>(raw[i]);
    else for (size_t i = 0; i < count; ++i) r.f[i] = static_cast<float>(static_cast<int8_t>(raw[i]));
  }
  if (!in) die("cannot read file " + file.string());
  return r;
}

// ------------------------------------------------------------------------------------------------ JSON (nlohmann dump(2) look-alike)
struct Json;
using JsonPtr = std::shared_ptr<Json>;
struct Json {
  std::variant<std::monostate, uint64_t, int64_t, double, std::string, std::map<std::string, JsonPtr>> v;
  Json& operator[](const std::string& k) {
    if (!std::holds_alternative<std::map<std::string, JsonPtr>>(v)) v = std::map<std::string, JsonPtr>{};
    auto& m = std::get<std::map<std::string, JsonPtr>>(v);
    auto it = m.find(k);
    if (it == m.end()) it = m.emplace(k, std::make_shared<Json>()).first;
    return *it->second;
  }
  template <class T>
  Json& operator=(T x) {
    if constexpr (std::is_same_v<T, bool>) v = std::string(x ? "true" : "false");  // the reference stores "true"/"false" strings
    else if constexpr (std::is_floating_point_v<T>) v = static_cast<double>(x);
    else if constexpr (std::is_integral_v<T> && std::is_signed_v<T>) v = static_cast<int64_t>(x);
    else if constexpr (std::is_integral_v<T>) v = static_cast<uint64_t>(x);
    else v = std::string(x);
    return *this;
  }
  static std::string esc(const std::string& s) {
    std::string o = "\"";
    for (char ch : s) {
      if (ch == '"' || ch == '\\') { o += '\\'; o += ch; }
      else if (ch == '\n') o += "\\n";
      else if (static_cast<unsigned char>(ch) < 0x20) { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", ch); o += b; }
      else o += ch;
    }
    return o + "\"";
  }
  void dump(std::ostream& os, int indent) const {
    if (auto p = std::get_if<uint64_t>(&v)) os << *p;
    else if (auto q = std::get_if<int64_t>(&v)) os << *q;
    else if (auto d = std::get_if<double>(&v)) {
      if (!std::isfinite(*d)) { os << "null"; return; }  // nlohmann prints NaN (e.g. a 0/0 hit rate) as null
      char buf[64];
      auto res = std::to_chars(buf, buf + sizeof buf, *d);
      std::string s(buf, res.ptr);
      if (s.find_first_of(".eE") == std::string::npos) s += ".0";
      os << s;
    } else if (auto s = std::get_if<std::string>(&v)) os << esc(*s);
    else if (auto m = std::get_if<std::map<std::string, JsonPtr>>(&v)) {
      if (m->empty()) { os << "{}"; return; }
      os << "{\n";
      size_t i = 0;
      for (auto& [k, child] : *m) {
        os << std::string(indent + 2, ' ') << esc(k) << ": ";
        child->dump(os, indent + 2);
        os << (++i < m->size() ? ",\n" : "\n");
      }
      os << std::string(indent, ' ') << "}";
    } else os << "null";
  }
};

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

fs::path find_stem(const fs::path& dir, const std::string& stem) {
  std::error_code ec;
  for (const auto& f : fs::directory_iterator(dir, ec)) {
    if (f.path().stem() == stem) return f.path();
  }
  return {};
}

// HNSW::estimate_index_size (src/hnsw/hnsw.hh:309-322)
uint64_t estimate_index_size(uint64_t num_nodes, uint32_t m, uint32_t dim) {
  uint64_t index_size = 0;
  const uint32_t num_levels = static_cast<uint32_t>(std::round(std::log(static_cast<double>(num_nodes)) / std::log(static_cast<double>(m))));
  const uint64_t node = 16 + 4ull * dim, l0 = 4 + 8ull * 2 * m, lu = 4 + 8ull * m;
  for (uint32_t i = 0; i < num_levels; ++i) {
    const uint64_t size = i == 0 ? node + l0 : lu;
    index_size += static_cast<uint64_t>(std::round(std::pow(1. / m, i) * static_cast<double>(num_nodes))) * size;
  }
  return index_size;
}

#define SHN(call) do { if ((call) != SHN_OK) die(std::string(#call) + ": " + shn_last_error()); } while (0)

}  // namespace

int main(int argc, char** argv) {
  const Config cfg = parse(argc, argv);
  if (cfg.is_server) {
    status("memory-node role: nothing to do, the index is resident in GPU memory of the compute node");
    return 0;
  }
  const uint32_t num_servers = static_cast<uint32_t>(cfg.servers.size());
  const bool compute_recall = !cfg.no_recall;
  if (cfg.ef_search < cfg.k) die("ef_search must be >= k");  // hnsw.hh:36

  // -------- read_dataset (compute_node.cc:278-320)
  const fs::path data_path = cfg.data_path;
  const fs::path base_file = find_stem(data_path, "base");
  const fs::path query_dir = data_path / "queries";
  const fs::path query_file = find_stem(query_dir, "query-" + cfg.query_suffix);
  const fs::path gt_file = find_stem(query_dir, "groundtruth-" + cfg.query_suffix);
  const fs::path warmup_file = find_stem(query_dir, "warmup-" + cfg.query_suffix);
  if (base_file.empty() || query_file.empty()) die("base or query file missing");
  Rows base = read_rows(base_file, /*meta_only=*/cfg.load_index);
  Rows queries = read_rows(query_file);
  if (queries.dim != base.dim) die("query dimension differs from base dimension");
  Rows warmup, gt;
  if (cfg.use_cache) {
    if (warmup_file.empty()) die("warmup file missing");
    warmup = read_rows(warmup_file);
  }
  if (compute_recall) {
    if (gt_file.empty()) die("ground truth file missing");
    gt = read_rows(gt_file);
    if (gt.dim < cfg.k || gt.n < queries.n) die("ground truth file is too small for k / the query file");
  }

  Json out;
  const shn_metric metric = cfg.ip_dist ? SHN_IP : SHN_L2;
  out["estimated_total_index_size"] = estimate_index_size(base.n, cfg.m, base.dim);
  out["distance"] = cfg.ip_dist ? "inner_product" : "squared_l2";
  out["node_size"] = uint64_t{16} + 4ull * base.dim;
  out["neighborlist_size"] = uint64_t{4} + 8ull * cfg.m;
  out["neighborlist_size_l0"] = uint64_t{4} + 16ull * cfg.m;
  out["cache"]["num_cache_buckets"] = uint64_t{0};
  out["cache"]["num_cooling_table_buckets"] = uint64_t{0};
  if (cfg.use_cache) out["cache"]["cache_size_ratio"] = cfg.cache_ratio;

  // -------- build or load (compute_node.cc:76-107)
  std::vector<std::string> dump_files;
  for (uint32_t i = 1; i <= num_servers; ++i) {
    dump_files.push_back((data_path / "dump" / ("index_m" + std::to_string(cfg.m) + "_efc" + std::to_string(cfg.ef_construction) +
                                                "_node" + std::to_string(i) + "_of" + std::to_string(num_servers) + ".dat")).string());
  }
  std::vector<const char*> dump_paths;
  for (auto& f : dump_files) dump_paths.push_back(f.c_str());

  if (cfg.gpus < 1 || cfg.gpus > 8) die("--gpus must be in [1, 8]");
  const int n_gpus = cfg.gpus;
  double build_ms = 0.0;
  shn_stats bstats;
  std::memset(&bstats, 0, sizeof bstats);
  uint64_t index_dump_bytes = 0, hbm_bytes = 0;
  uint32_t index_max_level = 0;
  std::vector<uint32_t> ids(static_cast<size_t>(std::max(queries.n, warmup.n)) * cfg.k);

  // One full index per GPU (the build is deterministic, so the graphs are identical).  With one GPU it is searched as
  // it is; with several, every GPU keeps its share (shn_index_partition) and reads the rest from its peers — the
  // reference's memory nodes become HBM partitions, its compute-node cache the replicated hot set.
  // one seed for every GPU's build: the graphs must be the same graph (with --seed -1 the clock is read once)
  const uint32_t build_seed = cfg.seed == -1 ? static_cast<uint32_t>(std::time(nullptr)) : static_cast<uint32_t>(cfg.seed);
  auto make_full = [&](int gpu, bool first) -> shn_index* {
    shn_index* full = nullptr;
    if (cfg.load_index) {
      if (first) status("load index from " + dump_files.front());
      SHN(shn_index_load(&full, dump_paths.data(), static_cast<int>(num_servers), base.dim, cfg.m, metric, gpu));
    } else {
      if (first) status("build index");
      const double t0 = now_ms();
      SHN(shn_index_build(&full, base.f.data(), nullptr, base.n, base.dim, cfg.m, cfg.ef_construction, metric, build_seed, gpu));
      if (first) {
        build_ms = now_ms() - t0;
        SHN(shn_index_build_stats(full, &bstats));
        status("processed inserts: " + std::to_string(base.n));
        if (cfg.store_index) {
          status("store index to " + dump_files.front());
          std::error_code ec;
          fs::create_directories(data_path / "dump", ec);
          SHN(shn_index_store(full, dump_paths.data(), static_cast<int>(num_servers)));
        }
      }
    }
    if (shn_index_size(full) != base.n) die("index holds " + std::to_string(shn_index_size(full)) + " nodes but base has " + std::to_string(base.n));
    if (first) { index_dump_bytes = shn_index_dump_bytes(full); index_max_level = shn_index_max_level(full); }
    return full;
  };

  std::vector<shn_index*> handles(n_gpus, nullptr);  // one GPU: the index; several: the full indexes until the group holds its partitions
  shn_group* group = nullptr;
  double placement_kmeans_ms = 0.0, placement_fetch_ms = 0.0, warmup_routing_ms = 0.0;
  for (int g = 0; g < n_gpus; ++g) {
    shn_index* full = make_full(cfg.gpu + g, g == 0);
    if (cfg.use_cache && warmup.n) {
      // cache warm-up pass (compute_node.cc:116-131), results discarded.  Several GPUs: the pass counts visits and the
      // most visited nodes join the replicated hot set — same queries on every GPU, so the same set everywhere
      if (g == 0) status("run warmup queries");
      if (n_gpus > 1) SHN(shn_index_count_visits(full, 1));
      SHN(shn_search(full, warmup.f.data(), warmup.n, cfg.k, cfg.ef_search, ids.data(), nullptr, nullptr));
    }
    handles[g] = full;
  }
  if (n_gpus == 1) {
    hbm_bytes = shn_index_hbm_bytes(handles[0]);
  } else {
    // memory nodes -> HBM partitions (NVLink peer reads), compute-node cache -> replicated hot set; with --routing the
    // nodes are placed by k-means cluster and the queries run on the GPU of their nearest centroid (compute_node.cc:110-131,
    // query_router.hh:280-387)
    // the cache budget (--cache-ratio, % of the nodes per compute node): without routing all of it is the replicated hot set;
    // with routing half of it, the other half is each GPU's own halo (the peer-owned rows its routed queries read most)
    const uint32_t cache_pct = cfg.use_cache ? cfg.cache_ratio : 0;
    const uint32_t halo_pct = (cfg.routing && warmup.n) ? cache_pct / 2 : 0;
    SHN(shn_group_create(&group, handles.data(), n_gpus, cache_pct - halo_pct, cfg.routing ? 1 : 0, cfg.routing ? 1 : 0,
                         0.25, std::max<uint64_t>(1, std::max(queries.n, warmup.n)), cfg.k, 1234));
    for (int g = 0; g < n_gpus; ++g) { shn_index_free(handles[g]); handles[g] = nullptr; }
    if (halo_pct) {
      const double w0 = now_ms();
      SHN(shn_group_warmup(group, warmup.f.data(), warmup.n, cfg.k, cfg.ef_search, halo_pct, nullptr));
      warmup_routing_ms = now_ms() - w0;
    }
    for (int g = 0; g < n_gpus; ++g) hbm_bytes += shn_index_hbm_bytes(shn_group_partition(group, g));
    SHN(shn_group_timings(group, &placement_kmeans_ms, &placement_fetch_ms));
  }
  out["allocated_local_buffer_size"] = hbm_bytes;
  out["actual_total_local_buffer_size"] = hbm_bytes;
  out["build"]["dist_comps"] = bstats.distcomps;
  out["build"]["rdma_reads_in_bytes"] = uint64_t{0};
  out["build"]["rdma_writes_in_bytes"] = uint64_t{0};
  out["build"]["remote_allocations"] = cfg.load_index ? uint64_t{0} : uint64_t{base.n};
  out["build"]["index_size"] = cfg.load_index ? uint64_t{0} : index_dump_bytes;
  out["build"]["max_level"] = cfg.load_index ? 0u : index_max_level;

  // -------- queries (compute_node.cc:140, 354-386): GPU g takes the queries with id % gpus == g (io/read_data.hh:58)
  status("run queries");
  shn_stats st;
  std::memset(&st, 0, sizeof st);
  std::vector<shn_stats> per_gpu(n_gpus);
  const double q0 = now_ms();
  double routing_ms = 0.0;
  if (n_gpus == 1) {
    SHN(shn_search(handles[0], queries.f.data(), queries.n, cfg.k, cfg.ef_search, ids.data(), nullptr, &st));
    per_gpu[0] = st;
  } else {
    // GPU g takes the queries with id % gpus == g (io/read_data.hh:58); the fan-out lives behind the C ABI (csrc/group.cu)
    SHN(shn_group_search(group, queries.f.data(), queries.n, cfg.k, cfg.ef_search, ids.data(), nullptr, per_gpu.data(), &routing_ms));
    for (int g = 0; g < n_gpus; ++g) {
      const shn_stats& p = per_gpu[g];
      st.distcomps += p.distcomps; st.visited_nodes += p.visited_nodes; st.visited_nodes_l0 += p.visited_nodes_l0;
      st.visited_neighborlists += p.visited_neighborlists; st.reference_layout_bytes += p.reference_layout_bytes;
      st.algorithmic_bytes += p.algorithmic_bytes; st.processed += p.processed;
      st.rows_hot += p.rows_hot + p.rows_halo; st.rows_local += p.rows_local; st.rows_remote += p.rows_remote;
      st.kernel_ms = std::max(st.kernel_ms, p.kernel_ms); st.h2d_ms = std::max(st.h2d_ms, p.h2d_ms); st.d2h_ms = std::max(st.d2h_ms, p.d2h_ms);
    }
  }
  const double query_ms = now_ms() - q0;
  status("processed queries: " + std::to_string(st.processed));

  // -------- compute_local_recall (compute_node.cc:579-600): query slot s has id s (single compute node)
  double recall = 0.0;
  if (compute_recall) {
    uint64_t hits = 0;
    for (uint32_t q = 0; q < queries.n; ++q) {
      const uint32_t* truth = gt.u.data() + static_cast<size_t>(q) * gt.dim;
      for (uint32_t j = 0; j < cfg.k; ++j) {
        const uint32_t hit = ids[static_cast<size_t>(q) * cfg.k + j];
        if (hit == 0xFFFFFFFFu) continue;
        for (uint32_t t = 0; t < cfg.k; ++t) {
          if (truth[t] == hit) { ++hits; break; }
        }
      }
    }
    recall = static_cast<double>(hits) / static_cast<double>(queries.n) / cfg.k;
    status("local recall: " + std::to_string(recall));
  }

  // -------- statistics (compute_node.cc:157-187, 478-558; statistics.hh:117-142)
  auto& qj = out["queries"];
  qj["dist_comps"] = st.distcomps;
  qj["rdma_reads_in_bytes"] = st.reference_layout_bytes;  // what the reference would have READ for the same traversal
  qj["rdma_writes_in_bytes"] = uint64_t{0};
  qj["recall"] = recall;
  qj["visited_nodes"] = st.visited_nodes;
  qj["visited_nodes_l0"] = st.visited_nodes_l0;
  qj["visited_neighborlists"] = st.visited_neighborlists;
  qj["processed"] = st.processed;
  for (int g = 0; g < n_gpus; ++g) qj["processed_local"]["c" + std::to_string(g)] = per_gpu[g].processed;
  qj["queries_per_sec"] = static_cast<uint64_t>(queries.n / (query_ms / 1000.0));
  qj["compute_recall"] = compute_recall;
  // extension keys (not in the reference): device-side view of the same run
  qj["gpu_kernel_ms"] = st.kernel_ms;
  qj["gpu_h2d_ms"] = st.h2d_ms;
  qj["gpu_d2h_ms"] = st.d2h_ms;
  qj["algorithmic_bytes"] = st.algorithmic_bytes;
  // several GPUs: a level-0 read served from the replicated hot set or the GPU's own share is a hit, a read from a
  // peer's share over NVLink a miss (what the reference would READ over RDMA)
  auto rate = [](uint64_t hits, uint64_t misses) { return static_cast<double>(hits) / static_cast<double>(hits + misses); };
  out["cache"]["hits_total"] = st.rows_hot + st.rows_local;
  out["cache"]["misses_total"] = st.rows_remote;
  out["cache"]["hit_rate"] = rate(st.rows_hot + st.rows_local, st.rows_remote);
  for (int g = 0; g < n_gpus; ++g)
    out["cache"]["local_hit_rates"]["c" + std::to_string(g)] = rate(per_gpu[g].rows_hot + per_gpu[g].rows_halo + per_gpu[g].rows_local, per_gpu[g].rows_remote);
  out["cache"]["local_size"] = uint64_t{0};
  out["cache"]["cached_nodes"] = uint64_t{0};
  out["cache"]["cache_buckets_size"] = uint64_t{0};

  const std::string path_name = data_path.has_stem() ? data_path.stem().string() : data_path.parent_path().stem().string();
  const size_t dash = cfg.query_suffix.find_first_of('-');
  const std::string zipf = cfg.query_suffix.size() > 1 ? cfg.query_suffix.substr(1, dash == std::string::npos ? std::string::npos : dash - 1) : "";
  auto& meta = out["meta"];
  meta["compute_nodes"] = static_cast<uint32_t>(n_gpus);
  meta["memory_nodes"] = num_servers;
  meta["compute_threads"] = cfg.threads;
  meta["coroutines_per_thread"] = cfg.coroutines;
  meta["threads_pinned"] = !cfg.disable_thread_pinning;
  meta["hyperthreading"] = false;
  meta["dataset"] = path_name;
  meta["query_suffix"] = cfg.query_suffix;
  meta["zipf_parameter"] = zipf;
  {
    const std::time_t now = std::time(nullptr);
    char buf[64];
    std::strftime(buf, sizeof buf, "%Y-%m-%dT%H:%M:%SZ", std::localtime(&now));  // timing.cc:105-112
    meta["timestamp"]["$date"] = std::string(buf);
  }
  meta["label"] = cfg.label;
  meta["engine"] = std::string(shn_version());
  out["hnsw_parameters"]["k"] = cfg.k;
  out["hnsw_parameters"]["m"] = cfg.m;
  out["hnsw_parameters"]["ef_search"] = cfg.ef_search;
  out["hnsw_parameters"]["ef_construction"] = cfg.ef_construction;
  out["num_vectors"] = base.n;
  out["num_queries"] = queries.n;
  auto& tj = out["timings"];
  tj["build_c0"] = build_ms;
  tj["query_c0"] = query_ms;
  for (int g = 1; g < n_gpus; ++g) { tj["build_c" + std::to_string(g)] = build_ms; tj["query_c" + std::to_string(g)] = query_ms; }
  tj["build_max"] = build_ms;
  tj["query_max"] = query_ms;
  tj["placement_fetch"] = placement_fetch_ms;    // several GPUs: splitting the index into the GPUs' shares
  tj["placement_kmeans"] = placement_kmeans_ms;  // --routing: k-means over the upper-level nodes + balanced assignment
  tj["routing"] = routing_ms;                    // --routing: route + deliver the batch (slowest GPU)
  if (cfg.use_cache) tj["warmup_routing"] = warmup_routing_ms;  // --routing: the routed warm-up pass that fills the GPUs' halos

  std::cerr << std::endl << "statistics:" << std::endl;
  out.dump(std::cout, 0);
  std::cout << std::endl;
  shn_group_free(group);
  for (shn_index* h : handles) shn_index_free(h);
  return 0;
}
