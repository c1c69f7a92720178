"""The compute-node host binary (dm-hnsw-reference_b200/shine_b200): the reference's CLI validation
(common/configuration.hh:88-113, rdma-library/library/configuration.cc:63-83), dataset directory, dump files and JSON
key tree (compute_node.cc:478-558, statistics.hh:117-142)."""
import json
import os
import subprocess

import numpy as np
import pytest

import datagen
import hnsw_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "dm-hnsw-reference_b200", "shine_b200")


def run(*args):
    return subprocess.run([BIN, *args], capture_output=True, text=True)


def test_binary_exists():
    assert os.path.exists(BIN), "run __graft_entry__.build()"


@pytest.mark.parametrize("args,msg", [
    (["--threads", "1"], "--servers <arg-list> must be given"),
    (["--is-server", "--initiator"], "a server cannot be the initiator"),
    (["--servers", "a", "--clients", "b"], "--clients <arg-list> is only required by the initiating client"),
    (["--servers", "a", "--threads", "1", "--ef-search", "10", "-k", "5"], "Data path and query suffix cannot be empty"),
    (["--servers", "a", "-d", "/tmp", "-q", "a0.0", "--ef-search", "10"], "Parameters threads, ef-search, and k are required"),
    (["--servers", "a", "-d", "/tmp", "-q", "a0.0", "-t", "1", "--ef-search", "10", "-k", "5", "--store-index", "--load-index"],
     "cannot be used in conjunction"),
    (["--servers", "a", "-d", "/tmp", "-q", "a0.0", "-t", "1", "--ef-search", "10", "-k", "5", "--routing"],
     "--routing can only be used in conjunction with --cache"),
    (["--servers", "a", "-d", "/tmp", "-q", "a0.0", "-t", "1", "--ef-search", "10", "-k", "5", "--cache", "--cache-ratio", "0"],
     "--cache-ratio must be > 0"),
    (["--servers", "a", "-d", "/tmp", "-q", "a0.0", "-t", "1", "--ef-search", "4", "-k", "5"], "ef_search must be >= k"),
    (["--servers", "a", "--bogus"], "unrecognised option"),
])
def test_cli_validation_matches_reference(args, msg):
    r = run(*args)
    assert r.returncode == 1 and msg in r.stderr and r.stdout == ""


def test_memory_node_role_is_a_no_op():
    r = run("--is-server", "--num-clients", "2", "--port", "1234")
    assert r.returncode == 0 and r.stdout == ""


def test_missing_dataset_fails_loudly(tmp_path):
    (tmp_path / "queries").mkdir()
    r = run("--servers", "a", "--initiator", "-d", str(tmp_path), "-q", "a0.0", "-t", "1", "--ef-search", "10", "-k", "5")
    assert r.returncode == 1 and "base or query file missing" in r.stderr


def test_argv_of_the_reference_launcher_is_accepted(tmp_path):
    """scripts/run_node.py:36-67 composes this command line for the compute-node role (benchmark.py sets config.EXECUTABLE);
    every flag must parse — the run then stops at the missing dataset, not at an option.  What the reference's fetch_*.py
    scripts read from the JSON (meta.zipf_parameter / label / dataset / compute_threads / compute_nodes,
    queries.queries_per_sec, cache.cache_size_ratio / hit_rate) is asserted on real runs in the GPU tests below."""
    (tmp_path / "queries").mkdir()
    argv = ["--servers", "cluster16", "cluster17", "--threads", "16", "--coroutines", "4", "--data-path", str(tmp_path) + "/",
            "--query-suffix", "a1.0-500k", "--ef-search", "100", "--ef-construction", "500", "--m", "32", "--k", "10",
            "--load-index", "--label", "routing", "--no-recall", "--ip-dist", "--cache", "--cache-ratio", "5", "--routing",
            "--initiator", "--clients", "cluster12", "cluster13"]
    r = run(*argv)
    assert r.returncode == 1 and "base or query file missing" in r.stderr, r.stderr


def write_bin(path, arr):
    with open(path, "wb") as f:
        np.array(arr.shape, dtype=np.uint32).tofile(f)
        arr.tofile(f)


REQUIRED_KEYS = {
    "": ["estimated_total_index_size", "allocated_local_buffer_size", "actual_total_local_buffer_size", "distance", "node_size",
         "neighborlist_size", "neighborlist_size_l0", "num_vectors", "num_queries", "meta", "hnsw_parameters", "build",
         "queries", "cache", "timings"],
    "meta": ["compute_nodes", "memory_nodes", "compute_threads", "coroutines_per_thread", "threads_pinned", "hyperthreading",
             "dataset", "query_suffix", "zipf_parameter", "timestamp", "label"],
    "hnsw_parameters": ["k", "m", "ef_search", "ef_construction"],
    "build": ["dist_comps", "rdma_reads_in_bytes", "rdma_writes_in_bytes", "remote_allocations", "index_size", "max_level"],
    "queries": ["dist_comps", "rdma_reads_in_bytes", "rdma_writes_in_bytes", "recall", "visited_nodes", "visited_nodes_l0",
                "visited_neighborlists", "processed", "queries_per_sec", "processed_local", "compute_recall"],
    "cache": ["hits_total", "misses_total", "hit_rate", "local_hit_rates", "num_cache_buckets", "num_cooling_table_buckets"],
    "timings": ["build_c0", "query_c0", "build_max", "query_max", "placement_fetch", "placement_kmeans", "routing"],
}


@pytest.mark.gpu
@pytest.mark.parametrize("ext,ip", [("fbin", False), ("u8bin", False), ("fbin", True)])
def test_build_store_load_query_round_trip(tmp_path, ext, ip):
    n, nq, dim, m, efc, k, ef = 5000, 300, 32, 16, 100, 10, 64
    base, queries = datagen.base_and_queries(n, nq, dim, normalize=ip)
    if ext == "u8bin":
        base = np.clip(np.round(base * 40 + 128), 0, 255).astype(np.uint8)
        queries = np.clip(np.round(queries * 40 + 128), 0, 255).astype(np.uint8)
    data = tmp_path / "synth-5k"
    (data / "queries").mkdir(parents=True)
    write_bin(data / f"base.{ext}", base)
    write_bin(data / "queries" / f"query-a0.0.{ext}", queries)
    gt = datagen.bruteforce(base.astype(np.float32), queries.astype(np.float32), 100, ip=ip)
    write_bin(data / "queries" / "groundtruth-a0.0.bin", gt)
    common = ["--servers", "mn1", "mn2", "--initiator", "--data-path", str(data), "--query-suffix", "a0.0", "--threads", "4",
              "--ef-search", str(ef), "--ef-construction", str(efc), "-k", str(k), "-m", str(m), "--label", "t"] + (["--ip-dist"] if ip else [])
    r = run(*common, "--store-index")
    assert r.returncode == 0, r.stderr
    doc = json.loads(r.stdout)  # scripts/benchmark.py:71 parses stdout as one JSON document
    for group, keys in REQUIRED_KEYS.items():
        node = doc[group] if group else doc
        for key in keys:
            assert key in node, f"{group}.{key}"
    assert doc["meta"]["dataset"] == "synth-5k" and doc["meta"]["zipf_parameter"] == "0.0" and doc["meta"]["memory_nodes"] == 2
    assert doc["distance"] == ("inner_product" if ip else "squared_l2")
    assert doc["node_size"] == 16 + 4 * dim and doc["neighborlist_size_l0"] == 4 + 16 * m and doc["neighborlist_size"] == 4 + 8 * m
    assert doc["num_vectors"] == n and doc["num_queries"] == nq and doc["queries"]["processed"] == nq
    assert doc["queries"]["recall"] > 0.95 and doc["build"]["dist_comps"] > 0 and doc["queries"]["queries_per_sec"] > 0
    assert doc["queries"]["compute_recall"] == "true" and doc["meta"]["threads_pinned"] == "true"
    dumps = [open(data / "dump" / f"index_m{m}_efc{efc}_node{i}_of2.dat", "rb").read() for i in (1, 2)]
    assert doc["build"]["index_size"] == sum(len(d) - 16 for d in dumps)

    # the dumps are reference-format: the oracle searches them and reproduces the binary's counters exactly
    oracle = hnsw_oracle.Index(dumps, dim, m)
    oi, _, _, ct = oracle.knn(queries.astype(np.float32), k, ef, ip=ip, counters=True, track_ties=True)
    if (ct["tie"] == 0).all():
        assert doc["queries"]["dist_comps"] == int(ct["distcomps"].sum())
        assert doc["queries"]["visited_neighborlists"] == int((ct["lists_l0"] + ct["lists_upper"]).sum())
        assert doc["queries"]["rdma_reads_in_bytes"] == int(ct["rdma_reads_in_bytes"].sum())
        assert abs(doc["queries"]["recall"] - datagen.recall(oi, gt[:, :k])) < 1e-9

    # --load-index on the same directory: same results, no build
    r2 = run(*common, "--load-index")
    assert r2.returncode == 0, r2.stderr
    doc2 = json.loads(r2.stdout)
    assert doc2["queries"]["recall"] == doc["queries"]["recall"] and doc2["queries"]["dist_comps"] == doc["queries"]["dist_comps"]
    assert doc2["build"]["dist_comps"] == 0 and doc2["timings"]["build_c0"] == 0.0


@pytest.mark.gpu
def test_two_gpus_one_partitioned_index(tmp_path):
    """--gpus 2: memory nodes become HBM partitions (in-process peer pointers), the warm-up pass picks the hot set; the
    answers must be those of the single-GPU run, and the cache keys of the JSON now carry real hit/miss counts."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n, nq, dim, m, efc, k, ef = 8000, 400, 32, 16, 100, 10, 64
    base, queries = datagen.base_and_queries(n, nq + 200, dim)
    data = tmp_path / "synth-8k"
    (data / "queries").mkdir(parents=True)
    write_bin(data / "base.fbin", base)
    write_bin(data / "queries" / "query-a0.0.fbin", queries[:nq])
    write_bin(data / "queries" / "warmup-a0.0.fbin", queries[nq:])
    write_bin(data / "queries" / "groundtruth-a0.0.bin", datagen.bruteforce(base, queries[:nq], 100))
    common = ["--servers", "mn1", "--initiator", "--data-path", str(data), "--query-suffix", "a0.0", "--threads", "4",
              "--ef-search", str(ef), "--ef-construction", str(efc), "-k", str(k), "-m", str(m)]
    one = run(*common, "--store-index")
    assert one.returncode == 0, one.stderr
    two = run(*common, "--load-index", "--gpus", "2", "--cache", "--cache-ratio", "10")
    assert two.returncode == 0, two.stderr
    d1, d2 = json.loads(one.stdout), json.loads(two.stdout)
    assert d2["queries"]["recall"] == d1["queries"]["recall"]
    assert d2["queries"]["dist_comps"] == d1["queries"]["dist_comps"]
    assert d2["meta"]["compute_nodes"] == 2 and set(d2["queries"]["processed_local"]) == {"c0", "c1"}
    assert d2["queries"]["processed_local"]["c0"] + d2["queries"]["processed_local"]["c1"] == nq
    c = d2["cache"]
    assert c["hits_total"] > 0 and c["misses_total"] > 0 and 0.5 < c["hit_rate"] < 1.0
    assert set(c["local_hit_rates"]) == {"c0", "c1"} and c["cache_size_ratio"] == 10
    assert d2["timings"]["placement_fetch"] > 0 and d2["timings"]["routing"] == 0.0
    # --routing (compute_node.cc:191-245): nodes placed by k-means cluster, queries answered by the GPU of their nearest
    # centroid; same answers, fewer remote reads, real timings in the reference's keys
    three = run(*common, "--load-index", "--gpus", "2", "--cache", "--cache-ratio", "10", "--routing")
    assert three.returncode == 0, three.stderr
    d3 = json.loads(three.stdout)
    assert d3["queries"]["recall"] == d1["queries"]["recall"]
    assert d3["queries"]["dist_comps"] == d1["queries"]["dist_comps"]
    assert d3["queries"]["processed_local"]["c0"] + d3["queries"]["processed_local"]["c1"] == nq
    assert d3["timings"]["placement_kmeans"] > 0 and d3["timings"]["routing"] > 0
    assert d3["cache"]["misses_total"] < c["misses_total"], "routing must cut the remote reads"
