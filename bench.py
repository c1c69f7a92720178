#!/usr/bin/env python
"""bench.py — QPS of the batched HNSW search path at recall@10 >= 0.9 on a SIFT10M-shaped synthetic index.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path (shn_search_device / shn_search) over one batch of synthetic queries.  The index
is resident in HBM before the timed region; `value` is measured with the queries already in HBM (CUDA events on the
launching stream), `e2e` goes through the host-buffer C-ABI call (H2D of the queries and D2H of the results inside the
timed region).  With N > 1 every rank holds a replica of the index, takes its own shard of the queries (weak scaling)
and the per-GPU top-k lists are gathered with an NCCL all-gather inside the timed region.

--impl reference times the reference's own CPU search path (oracle/_ref/libshine_ref.so = /root/reference's hnsw.hh
compiled unmodified; falls back to the plain-C port when that file is absent) on the host cores, on a bounded sample
of the same workload.  The oracle is only ever executed in that arm and in the cpu_baseline leg.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import __graft_entry__ as ge  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]: SIFT10M-shaped synthetic, single B200, ef sweep 16-256
    "sift10m": dict(n=10_000_000, dim=128, ip=False, m=16, efc=200, normalize=False,
                    label="SIFT10M-shaped synthetic (10M x 128 fp32, L2, M=16, efC=200), ef sweep 16-256, k=10"),
    # BASELINE.json configs[0]
    "sift1m": dict(n=1_000_000, dim=128, ip=False, m=16, efc=200, normalize=False,
                   label="SIFT1M-shaped synthetic (1M x 128 fp32, L2, M=16, efC=200), ef sweep 16-256, k=10"),
    # BASELINE.json configs[4]: Text-to-Image-10M-shaped, inner product, skewed queries (--zipf)
    "t2i10m": dict(n=10_000_000, dim=200, ip=True, m=16, efc=200, normalize=True,
                   label="Text-to-Image-10M-shaped synthetic (10M x 200 fp32, inner product, M=16, efC=200), ef sweep 16-256, k=10"),
    # BASELINE.json configs[3]: GIST1M-shaped, the high-dimensional bandwidth-bound case
    "gist1m": dict(n=1_000_000, dim=960, ip=False, m=16, efc=200, normalize=False,
                   label="GIST1M-shaped synthetic (1M x 960 fp32, L2, M=16, efC=200), ef sweep 16-256, k=10"),
    # BASELINE.json configs[2]: DEEP100M-shaped
    "deep100m": dict(n=100_000_000, dim=96, ip=False, m=16, efc=200, normalize=True,
                     label="DEEP100M-shaped synthetic (100M x 96 fp32 unit rows, L2, M=16, efC=200), ef sweep 16-256, k=10"),
    "tiny": dict(n=100_000, dim=128, ip=False, m=16, efc=200, normalize=False,
                 label="100k x 128 fp32 synthetic (smoke-sized), L2, M=16, efC=200"),
}
EF_SWEEP = (16, 32, 64, 128, 256)
K = 10
RECALL_TARGET = 0.9


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def synth_rows(n, dim, seed, device, normalize=False, r=16, chunk=1 << 20):
    """SURVEY 8d generator: x = A z / sqrt(r) + 0.05 eps with a fixed A (seed 1), z ~ N(0, I_r)."""
    ga = torch.Generator(device="cpu").manual_seed(1)
    a = torch.randn(dim, r, generator=ga).to(device)
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((n, dim), dtype=torch.float32, device=device)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        z = torch.randn(e - s, r, generator=g, device=device)
        x = z @ a.T / (r ** 0.5) + 0.05 * torch.randn(e - s, dim, generator=g, device=device)
        if normalize:
            x = x / x.norm(dim=1, keepdim=True)
        out[s:e] = x
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.lines = gpu, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], 0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "samples": len(sm),
                "reasons": sorted(reasons)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes():
    """dram bytes per launch of the search kernel from the committed ncu summary, if there is one."""
    try:
        with open(os.path.join(ROOT, "profiles", "search_kernel_traffic.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return None


def build_index(pkg, wl, base_dev, gpu):
    """Index over base_dev in HBM, built on the GPU (shn_index_build_device) — never inside a timed region."""
    torch.cuda.synchronize()  # base_dev was written on torch's stream, the library builds on its own
    t0 = time.time()
    ix = pkg.Index.build_device(base_dev.data_ptr(), wl["n"], wl["dim"], wl["m"], wl["efc"], ip=wl["ip"], seed=1234, gpu=gpu)
    torch.cuda.synchronize()
    return ix, "gpu-built (shn_index_build_device)", time.time() - t0


def ground_truth(pkg, base_dev, q_dev, k, ip, gpu):
    """Exact top-k for the recall check: shn_bruteforce_topk_device (csrc/bruteforce.cu), outside every timed region."""
    nq = q_dev.shape[0]
    out = torch.empty((nq, k), dtype=torch.int32, device=q_dev.device)
    torch.cuda.synchronize()
    pkg.bruteforce_topk_device(base_dev.data_ptr(), base_dev.shape[0], q_dev.data_ptr(), nq, base_dev.shape[1], k,
                               out.data_ptr(), ip=ip, gpu=gpu)
    return out.long()


def recall_at_k(ids_dev, gt_dev):
    k = gt_dev.shape[1]
    hit = (ids_dev.long()[:, :, None] == gt_dev[:, None, :]).any(2).sum().item()
    return hit / (gt_dev.shape[0] * k)


def cpu_search(dumps, wl, queries_np, ef, threads, budget_s):
    """The reference's CPU search path on a bounded sample: oracle/_ref when present (kind 'reference'), else the
    plain-C port (kind 'port').  Returns (qps, kind, nq_done, seconds, ids [nq_done,K] u32, total distcomps)."""
    import shine_ref
    if shine_ref.available():
        probe = queries_np[:256]
        _, _, _, _, s = shine_ref.search(dumps, wl["dim"], wl["m"], probe, K, ef, ip=wl["ip"], threads=threads, coroutines=4)
        rate = len(probe) / max(s, 1e-6)
        nq = int(min(len(queries_np), max(512, rate * budget_s)))
        ids, _, _, st, s = shine_ref.search(dumps, wl["dim"], wl["m"], queries_np[:nq], K, ef, ip=wl["ip"], threads=threads, coroutines=4)
        return nq / s, "reference", nq, s, ids, int(st["distcomps"])
    import hnsw_oracle
    ix = hnsw_oracle.Index(dumps, wl["dim"], wl["m"])
    probe = queries_np[:256]
    t0 = time.time(); ix.knn(probe, K, ef, ip=wl["ip"], threads=threads); s = time.time() - t0
    nq = int(min(len(queries_np), max(512, len(probe) / max(s, 1e-6) * budget_s)))
    t0 = time.time(); ids, _, _, ct = ix.knn(queries_np[:nq], K, ef, ip=wl["ip"], threads=threads, counters=True); s = time.time() - t0
    return nq / s, "port", nq, s, ids, int(ct["distcomps"].sum())


def parity_at_scale(ix, q_dev, ef, cpu_ids, cpu_distcomps, stream):
    """The CPU arm's answers for its sample against the GPU's for the same queries on the same index (the dump the CPU arm
    searched was written from this very handle): fraction of queries with the same id set, and whether the total number
    of distance computations agrees (it does unless an exact distance tie was broken differently, DESIGN.md 'Ties')."""
    nq = cpu_ids.shape[0]
    ids = torch.empty((nq, K), dtype=torch.int32, device=q_dev.device)
    st = ix.search_device(q_dev[:nq].contiguous().data_ptr(), nq, K, ef, ids.data_ptr(), stream=stream)
    gpu = np.sort(ids.cpu().numpy().view(np.uint32), axis=1)
    same = (gpu == np.sort(cpu_ids, axis=1)).all(axis=1)
    return dict(queries=int(nq), id_identical_frac=round(float(same.mean()), 6),
                distcomps_equal=bool(st["distcomps"] == cpu_distcomps),
                distcomps_rel_diff=float(abs(int(st["distcomps"]) - int(cpu_distcomps)) / max(1, int(cpu_distcomps))),
                distcomps_gpu=int(st["distcomps"]), distcomps_cpu=int(cpu_distcomps))


CPU_ARM_FLAGS = "-O2 -march=x86-64-v3 -ffast-math (the reference's CMakeLists.txt says -march=native; the .so must also run on the GPU box's host)"


def exchange_fds(my_fd, my_size, rank, world, dist, tag):
    """Every rank hands one POSIX fd (a CUDA VMM export) to every other rank over Unix sockets (SCM_RIGHTS).  Returns
    {peer: (fd, size)}; the caller closes what it received and what it exported."""
    import socket
    import threading
    name = lambda r: f"\0shn-{tag}-{r}"
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(name(rank))
    srv.listen(world)

    def serve():
        for _ in range(world - 1):
            conn, _ = srv.accept()
            with conn:
                socket.send_fds(conn, [str(my_size).encode()], [my_fd])

    th = threading.Thread(target=serve, daemon=True)
    th.start()
    dist.barrier()
    got = {}
    for peer in range(world):
        if peer == rank:
            continue
        with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
            c.connect(name(peer))
            msg, fds, _, _ = socket.recv_fds(c, 64, 1)
        got[peer] = (fds[0], int(msg.decode()))
    th.join()
    srv.close()
    dist.barrier()
    return got


def partitioned_block(pkg, args, dist, rank, world, dev, local_rank, peak):
    """north_star's multi-GPU design, measured: ONE graph partitioned over the GPUs' HBM (nodes stored on the GPU of their
    nearest k-means centroid, src/cache/placement.hh), a replicated hot set as the compute-node cache (src/cache/cache.hh),
    remote hops as NVLink peer loads (src/rdma/rdma_reads.hh), queries routed to the GPU of their nearest centroid on the
    GPU and exchanged through peer-mapped inboxes (src/router/query_router.hh; csrc/router.cu), per-GPU top-k lists merged
    with an NCCL all-gather.  Returns the `partitioned` object of the JSON line (rank 0) or None."""
    wl = dict(WORKLOADS[args.part_workload])
    nq = args.queries_per_step
    par = pkg.parallel
    t_all = time.time()
    base = synth_rows(wl["n"], wl["dim"], 1001, dev, wl["normalize"])
    ix, how, build_s = build_index(pkg, wl, base, local_rank)
    log(f"[rank {rank}] partitioned block: {wl['label']}: {how} in {build_s:.1f}s, hbm={ix.hbm_bytes / 1e9:.1f} GB")
    tstream = torch.cuda.current_stream()
    stream = tstream.cuda_stream
    batches = [synth_rows(nq, wl["dim"], 2002 + 1000 * rank + b, dev, wl["normalize"]) for b in range(2)]
    ids = torch.empty((nq, K), dtype=torch.int32, device=dev)
    dists = torch.empty((nq, K), dtype=torch.float32, device=dev)

    # ground truth + ef (rank 0's batch; every rank computes its own so that all stay in step)
    nrec = min(args.part_recall_queries, nq)
    gt = ground_truth(pkg, base, batches[0][:nrec].contiguous(), K, wl["ip"], local_rank)
    del base
    torch.cuda.empty_cache()
    efs = [args.part_ef] if args.part_ef else [100]
    sweep = []
    ef = efs[0]
    whole_ids = None  # the whole-index answers for batch 0 at the block's ef: what the partitioned index must return too
    for e in sorted(set(efs + [64])):
        st = ix.search_device(batches[0].data_ptr(), nq, K, e, ids.data_ptr(), dists.data_ptr(), stream=stream)
        sweep.append(dict(ef=e, recall=round(recall_at_k(ids[:nrec], gt), 4), whole_index_qps_per_gpu=round(nq / st["kernel_ms"] * 1e3, 1),
                          alg_bytes_per_query=round(st["algorithmic_bytes"] / nq, 1)))
        log(f"[rank {rank}] whole index on one GPU: {sweep[-1]}")
        if e == ef:
            whole_ids = ids.clone()
    whole = next(s for s in sweep if s["ef"] == ef)

    # warm-up pass with visit counting (compute_node.cc:116-131) -> the same hot set on every rank
    t0 = time.time()
    warm = synth_rows(min(nq, 200_000), wl["dim"], 7007 + rank, dev, wl["normalize"])
    ix.count_visits(True)
    torch.cuda.synchronize()
    ix.search_device(warm.data_ptr(), warm.shape[0], K, ef, ids.data_ptr(), stream=stream)
    counts = torch.empty(ix.n, dtype=torch.int32, device=dev)
    ix.visit_counts(counts.data_ptr())
    dist.all_reduce(counts)
    torch.cuda.synchronize()   # the library copies on its own stream: the all-reduce must have finished
    ix.visit_counts(counts.data_ptr(), write_back=True)
    owner = torch.empty(ix.n, dtype=torch.uint8, device=dev)
    t1 = time.time()
    centroids, sizes = ix.placement_fit(world, owner.data_ptr(), seed=1234, slack=0.05)  # deterministic: same on every rank
    t_fit = time.time() - t1
    part = ix.partition(rank, world, args.cache_ratio, d_owner=owner.data_ptr())
    ix.close()
    del counts, warm, owner
    torch.cuda.empty_cache()
    par.exchange_partition_shares(part, rank, world, dist, "p" + os.environ.get("MASTER_PORT", "0"))
    router = pkg.Router(part, centroids, slack=args.route_slack, max_batch=nq, k_max=K)
    fd, size, _ = router.export(want_fd=True)
    for peer, (pfd, psize) in exchange_fds(fd, size, rank, world, dist, "r" + os.environ.get("MASTER_PORT", "0")).items():
        router.attach(peer, fd=pfd, size=psize)
        os.close(pfd)
    os.close(fd)
    dist.barrier()
    setup_s = time.time() - t0
    token = torch.zeros(1, dtype=torch.int32, device=dev)

    def fence():  # stream-ordered barrier between the ranks: every peer's preceding kernels have completed
        dist.all_reduce(token)

    # halo: a routed warm-up pass with visit counting, then every GPU caches the peer-owned rows it read most
    halo_rows = 0
    if args.halo_ratio:
        t_h = time.time()
        part.count_visits(True)
        warm2 = synth_rows(min(nq, 200_000), wl["dim"], 9009 + rank, dev, wl["normalize"])
        torch.cuda.synchronize(); dist.barrier()
        router.scatter(warm2.data_ptr(), warm2.shape[0], stream=stream)
        fence()
        router.search(K, ef, stream=stream, want_stats=False)
        fence()
        torch.cuda.synchronize(); dist.barrier()
        halo_rows = part.build_halo(args.halo_ratio)
        dist.barrier()
        log(f"[rank {rank}] halo: {halo_rows} rows ({100.0 * halo_rows / wl['n']:.2f}% of the nodes) in {time.time() - t_h:.1f}s")
        del warm2
    setup_s = time.time() - t0
    log(f"[rank {rank}] partitioned x{world}: part sizes {sizes.tolist()}, hbm {part.hbm_bytes / 1e9:.1f} GB, setup {setup_s:.1f}s (k-means fit {t_fit:.1f}s)")
    # every rank must have cut the same graph the same way (a share is addressed with the reader's numbering)
    p_hot, _, p_ep = part.partition_info()
    sig = torch.tensor([p_hot, p_ep] + [int(x) for x in sizes], dtype=torch.int64, device=dev)
    sigs = [torch.zeros_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig)
    if any(not torch.equal(sigs[0], x) for x in sigs):
        raise SystemExit(f"[rank {rank}] the ranks partitioned different graphs: {[x.tolist() for x in sigs]}")
    # parity of the partitioned index before anything is timed: the unrouted search of this rank's partition (hot set + own
    # share + NVLink peer reads) must return what the whole index returned
    part.search_device(batches[0].data_ptr(), nq, K, ef, ids.data_ptr(), stream=stream)
    part_identical = float((ids == whole_ids).all(1).float().mean())
    p_ids, p_d = router.results()
    land_i = pkg.device_view(p_ids, (nq, K), "<i4")
    land_d = pkg.device_view(p_d, (nq, K), "<f4")
    def step(i, events=None):
        q = batches[i % 2]
        router.scatter(q.data_ptr(), nq, stream=stream)
        fence()
        if events: events[0].record()
        router.search(K, ef, stream=stream, want_stats=False)
        fence()
        if events: events[1].record()
        return par.allgather_results(land_i, land_d, world * nq, rank, world, dist)

    for i in range(args.warmup):
        step(i)
    dist.barrier(); torch.cuda.synchronize()
    steps = max(2, min(args.steps, args.part_steps))
    clocks = ClockSampler(local_rank)
    clocks.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    for i in range(steps):
        step(args.warmup + i)
    ev[1].record()
    dist.barrier(); torch.cuda.synchronize()
    total_ms = ev[0].elapsed_time(ev[1])
    clk = clocks.stop()
    # one more step with the phases timed and the counters read (outside the timed region)
    dist.barrier(); torch.cuda.synchronize()
    fence()
    ev[0].record()
    router.scatter(batches[0].data_ptr(), nq, stream=stream)
    fence()
    ev[1].record()
    st = router.search(K, ef, stream=stream)
    fence()
    ev[2].record()
    g_ids, _ = par.allgather_results(land_i, land_d, world * nq, rank, world, dist)
    ev[3].record()
    torch.cuda.synchronize()
    sent, received = router.counts(stream=stream)
    rec = recall_at_k(land_i[:nrec], gt)
    routed_identical = float((land_i == whole_ids).all(1).float().mean())
    phase = dict(route_scatter_ms=round(ev[0].elapsed_time(ev[1]), 3), search_ms=round(ev[1].elapsed_time(ev[2]), 3),
                 search_kernel_ms=round(st["kernel_ms"], 3), allgather_ms=round(ev[2].elapsed_time(ev[3]), 3))
    tot = max(1, st["rows_hot"] + st["rows_local"] + st["rows_remote"] + st["rows_halo"])
    remote_bytes = st["rows_remote"] * (4 * wl["dim"]) + (st["rows_remote"] / tot) * st["lists_l0"] * 8 * wl["m"]
    t = torch.tensor([total_ms, phase["search_kernel_ms"], remote_bytes / max(1e-9, st["kernel_ms"] * 1e-3) / 1e9,
                      st["algorithmic_bytes"] / max(1e-9, st["kernel_ms"] * 1e-3) / 1e9], device=dev, dtype=torch.float64)
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    frac = torch.tensor([st["rows_hot"], st["rows_local"], st["rows_remote"], st["processed"], st["rows_halo"], halo_rows],
                        device=dev, dtype=torch.float64)
    dist.all_reduce(frac)
    ident = torch.tensor([part_identical, routed_identical], device=dev, dtype=torch.float64)
    dist.all_reduce(ident, op=dist.ReduceOp.MIN)
    router.close(); part.close()
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    total_ms = float(tmax[0])
    qps = world * nq * steps / (total_ms * 1e-3)
    rows = float(frac[0] + frac[1] + frac[2] + frac[4])
    return dict(workload=wl["label"], ef=ef, k=K, recall_at_10=round(rec, 4), recall_whole_index=whole["recall"],
                identical_to_whole_index=dict(partition_unrouted=round(float(ident[0]), 6), routed=round(float(ident[1]), 6),
                                              note="fraction of the step's queries whose id list equals the whole-index answer, min over ranks (the partitioned kernel "
                                                   "evaluates the NVLink rows of a list first: two candidates of one list at the exact same "
                                                   "distance can swap)"),
                processed_all_ranks=int(frac[3]), value=round(qps, 1), unit="queries/s",
                steps=steps, ms_per_step=round(total_ms / steps, 3), queries_per_step_per_gpu=nq,
                design="graph partitioned by k-means cluster over the GPUs' HBM + replicated hot set + per-GPU halo (local copies of the "
                       "peer-owned rows the GPU's routed queries read most); remote hops = NVLink peer loads; "
                       "queries routed on the GPU and delivered into peer-mapped inboxes, results written to the home GPU's landing "
                       "buffer by the search kernel; NCCL all-gather of the top-k lists",
                whole_index_qps_per_gpu=whole["whole_index_qps_per_gpu"],
                efficiency_vs_whole_index_replicas=round(qps / (world * whole["whole_index_qps_per_gpu"]), 4),
                rows_hot=round(float(frac[0]) / rows, 4), rows_local=round(float(frac[1]) / rows, 4), rows_remote=round(float(frac[2]) / rows, 4),
                rows_halo=round(float(frac[4]) / rows, 4), halo_ratio_pct=args.halo_ratio,
                halo_rows_per_gpu=int(float(frac[5]) / world),
                nvlink_in_gbs_max=round(float(tmax[2]), 1), nvlink_peak_gbs=770.0, nvlink_frac=round(float(tmax[2]) / 770.0, 4),
                hbm_alg_gbs_per_gpu=round(float(t[3]), 1), hbm_frac=round(float(t[3]) / peak, 4),
                cache_ratio_pct=args.cache_ratio, route_slack=args.route_slack, step_ms_rank0=phase,
                sent_rank0=sent.tolist(), received_rank0=received.tolist(), part_sizes=[int(x) for x in sizes],
                index_build_s=round(build_s, 1), setup_s=round(setup_s, 1), block_s=round(time.time() - t_all, 1),
                sweep_whole_index=sweep, clocks=clk)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SHN_BENCH_WORKLOAD", "sift10m"), choices=sorted(WORKLOADS))
    ap.add_argument("--queries-per-step", type=int, default=1_000_000)
    ap.add_argument("--recall-queries", type=int, default=10_000)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ef", type=int, default=0, help="fix ef instead of picking the smallest with recall >= 0.9")
    ap.add_argument("-m", type=int, default=0, help="override the workload's M (the reference's experiments use 32, scripts/config.py:5-9)")
    ap.add_argument("--ef-construction", type=int, default=0, help="override the workload's efC (the reference's experiments use 500)")
    ap.add_argument("--zipf", type=float, default=None,
                    help="skew the queries: a pool of 100k distinct queries expanded with Zipf(alpha) popularity exactly as the "
                         "reference's scripts/data/skew.py")
    # N > 1: besides the replica headline, the partitioned design north_star names is measured in the same run
    ap.add_argument("--part-workload", default=os.environ.get("SHN_BENCH_PART_WORKLOAD", "deep100m"), choices=sorted(WORKLOADS),
                    help="N > 1: the index that is partitioned over the GPUs (BASELINE.json configs[2] by default)")
    ap.add_argument("--part-ef", type=int, default=0, help="ef of the partitioned block (default 100, scripts/datasets.py:16)")
    ap.add_argument("--part-steps", type=int, default=5)
    ap.add_argument("--part-recall-queries", type=int, default=5_000)
    ap.add_argument("--cache-ratio", type=int, default=8, help="partitioned: replicated hot set in %% of the nodes (--cache-ratio of the reference)")
    ap.add_argument("--halo-ratio", type=int, default=16,
                    help="partitioned: every GPU also caches this %% of the nodes from its peers' shares — the rows its own routed "
                         "queries read most (shn_index_partition_build_halo); 0 = off")
    ap.add_argument("--route-slack", type=float, default=0.25, help="a GPU takes at most (1 + slack) / N of a batch (query_router.hh:106-151)")
    ap.add_argument("--partitioned", default="auto", choices=["auto", "off", "only"],
                    help="auto: measured when N > 1; off: replica headline only; only: skip the headline (development)")
    args = ap.parse_args()
    if args.warmup < 3:
        log("bench: raising --warmup to 3 (timing rules)")
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = dict(WORKLOADS[args.workload])
    if args.m:
        wl["m"] = args.m
        wl["label"] = wl["label"].replace("M=16", f"M={args.m}")
    if args.ef_construction:
        wl["efc"] = args.ef_construction
        wl["label"] = wl["label"].replace("efC=200", f"efC={args.ef_construction}")

    if args.impl == "reference" and rank != 0:
        return  # rank 0 alone runs the CPU arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU path (the reference arm builds its index on the GPU too)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1 and args.impl == "b200":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    pkg = ge.load_package()
    peak, peak_src = measured_peak()
    # a non-default torch stream: the C ABI launches on the stream it is handed, and torch.cuda.Event only sees
    # torch's current stream
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    if args.partitioned == "only":
        if dist is None:
            raise SystemExit("--partitioned only needs N > 1 (torchrun)")
        block = partitioned_block(pkg, args, dist, rank, world, dev, local_rank, peak)
        if rank == 0:
            print(json.dumps(dict(n_gpus=world, partitioned=block)), flush=True)
        dist.destroy_process_group()
        return

    nq = args.queries_per_step
    base = synth_rows(wl["n"], wl["dim"], 1001, dev, wl["normalize"])
    ix, how, build_s = build_index(pkg, wl, base, local_rank)
    log(f"[rank {rank}] index: {how}, {build_s:.1f}s, n={ix.n} max_level={ix.max_level} hbm={ix.hbm_bytes / 1e9:.2f} GB")

    # query batches: held-out draws of the same model; each rank its own shard (seed), 4 distinct batches rotate
    n_batches = 4
    def make_queries(count, seed):
        if args.zipf is None:
            return synth_rows(count, wl["dim"], seed, dev, wl["normalize"])
        import datagen
        pool = synth_rows(100_000, wl["dim"], 2002, dev, wl["normalize"])  # the same pool on every rank
        pick = torch.from_numpy(datagen.zipf_indices(100_000, count, args.zipf, seed=seed)).to(dev)
        return pool[pick].contiguous()

    batches = [make_queries(nq, 2002 + 1000 * rank + b) for b in range(n_batches)]
    ids = torch.empty((nq, K), dtype=torch.int32, device=dev)
    dists = torch.empty((nq, K), dtype=torch.float32, device=dev)

    # ---- ef selection: the smallest ef of the sweep with recall@10 >= 0.9 (rank 0 decides) -----------------------
    nrec = min(args.recall_queries, nq)
    t0 = time.time()
    gt = ground_truth(pkg, base, batches[0][:nrec].contiguous(), K, wl["ip"], local_rank)
    log(f"[rank {rank}] ground truth for {nrec} queries (brute force on the GPU): {time.time() - t0:.1f}s")
    del base
    sweep = []
    for ef in sorted(set(EF_SWEEP) | ({args.ef} if args.ef else set())):
        st = ix.search_device(batches[0].data_ptr(), nq, K, ef, ids.data_ptr(), dists.data_ptr(), stream=stream)
        rec = recall_at_k(ids[:nrec], gt)
        sweep.append(dict(ef=ef, recall=round(rec, 4), qps=round(nq / st["kernel_ms"] * 1e3, 1),
                          distcomps_per_query=round(st["distcomps"] / nq, 1),
                          alg_bytes_per_query=round(st["algorithmic_bytes"] / nq, 1),
                          hbm_frac=round(st["algorithmic_bytes"] / (st["kernel_ms"] * 1e-3) / 1e9 / peak, 4),
                          overflow_queries=st["overflow_queries"]))
        log(f"[rank {rank}] sweep {sweep[-1]}")
    ef = args.ef or next((s["ef"] for s in sweep if s["recall"] >= RECALL_TARGET), EF_SWEEP[-1])
    recall = next((s["recall"] for s in sweep if s["ef"] == ef), None)

    def make_config():
        """The same dict in both arms (the driver compares them key by key)."""
        return dict(workload=wl["label"], ef=ef, k=K, recall_at_10=recall, queries_per_step_per_gpu=nq,
                    index=how, zipf_alpha=args.zipf,
                    l2_policy="index (>= 6 GB at 10M rows) and the rotating query batches are larger than the 126 MB L2; no flush",
                    parallelism=(f"replica x{args.gpus}, queries sharded, NCCL all-gather of top-k" if args.gpus > 1 else "single GPU"))

    if args.impl == "reference":
        threads = os.cpu_count()
        dumps = [d.tobytes() for d in ix.to_dumps(1)]
        qnp = batches[0].cpu().numpy()
        per_step = []
        sample = None
        worst = None
        for i in range(args.warmup + args.steps):
            qps, kind, done, secs, cpu_ids, cpu_dc = cpu_search(dumps, wl, qnp, ef, threads, max(2.0, args.cpu_seconds / max(1, args.steps)))
            sample = f"{done} queries of the step's batch per step (ef={ef}, k={K}), {secs:.1f}s"
            if i >= args.warmup:
                per_step.append((done, secs))
                par_s = parity_at_scale(ix, batches[0], ef, cpu_ids, cpu_dc, stream)  # outside the CPU timing
                if worst is None or par_s["id_identical_frac"] < worst["id_identical_frac"]:
                    worst = par_s
        tot_q = sum(d for d, _ in per_step); tot_s = sum(s for _, s in per_step)
        val = tot_q / tot_s
        line = dict(metric="queries_per_sec at recall@10>=0.9", value=round(val, 1), unit="queries/s", impl="reference",
                    n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=round(1e3 * tot_s / len(per_step), 3),
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    config=make_config(),
                    cpu_baseline=dict(value=round(val, 1), unit="queries/s", cores=threads, kind=kind, sample=sample,
                                      flags=CPU_ARM_FLAGS),
                    e2e=dict(value=round(val, 1), unit="queries/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    parity_at_scale=worst)
        print(json.dumps(line), flush=True)
        if worst and worst["id_identical_frac"] < 0.99:
            raise SystemExit(f"parity at scale failed: {worst}")
        return

    # ---- device-resident timing (value) ------------------------------------------------------------------------
    par = pkg.parallel

    def step_device(i):
        q = batches[i % n_batches]
        return ix.search_device(q.data_ptr(), nq, K, ef, ids.data_ptr(), dists.data_ptr(), stream=stream, want_stats=False)

    def gather():  # per-GPU top-k lists -> every rank, global query order; one all-gather
        return par.allgather_results(ids, dists, world * nq, rank, world, dist)

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step_device(i)
        if dist:
            gather()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = []
    torch.cuda.profiler.start()  # no-op unless under ncu --profile-from-start off
    ev0.record()
    for i in range(args.steps):
        step_device(args.warmup + i)
        if dist:
            gather()
    ev1.record()
    barrier()
    torch.cuda.profiler.stop()
    total_ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()

    # kernel-only duration + algorithmic bytes of the dominant kernel, measured live (events inside the C ABI on the
    # launching stream), same batches
    alg_bytes = 0
    for i in range(args.steps):
        st = ix.search_device(batches[(args.warmup + i) % n_batches].data_ptr(), nq, K, ef, ids.data_ptr(), dists.data_ptr(),
                              stream=stream)
        kern_ms.append(st["kernel_ms"]); alg_bytes += st["algorithmic_bytes"]
    kern_avg_ms = float(np.mean(kern_ms))
    achieved = alg_bytes / args.steps / (kern_avg_ms * 1e-3) / 1e9
    # ---- end to end through the host-buffer C-ABI call (e2e) -------------------------------------------------------
    host_q = [torch.empty((nq, wl["dim"]), dtype=torch.float32).pin_memory() for _ in range(2)]
    for h, b in zip(host_q, batches):
        h.copy_(b)
    host_ids = torch.empty((nq, K), dtype=torch.int32).pin_memory()
    host_d = torch.empty((nq, K), dtype=torch.float32).pin_memory()

    def step_host(i):
        ix.search(host_q[i % 2].numpy(), K, ef, out_ids=host_ids.numpy().view(np.uint32), out_dists=host_d.numpy())

    for i in range(2):
        step_host(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_host(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    # max over ranks
    if dist:
        t = torch.tensor([total_ms, e2e_s * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = t.tolist()
        e2e_s = e2e_ms / 1e3
    value = world * nq * args.steps / (total_ms * 1e-3)
    e2e = world * nq * args.steps / e2e_s

    block = None
    if dist and args.partitioned == "auto":
        ix.close()
        del batches, ids, dists, host_q, host_ids, host_d, gt
        torch.cuda.empty_cache()
        block = partitioned_block(pkg, args, dist, rank, world, dev, local_rank, peak)

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    traffic = ncu_traffic_bytes()
    if traffic and not (traffic.get("workload") == args.workload and traffic.get("ef") == ef and
                        traffic.get("queries_per_launch") == nq):
        traffic = None  # the committed ncu capture is for another configuration
    line = dict(metric="queries_per_sec at recall@10>=0.9", value=round(value, 1), unit="queries/s", n_gpus=world,
                steps=args.steps, warmup=args.warmup, ms_per_step=round(total_ms / args.steps, 3), higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=make_config(),
                roofline=dict(bound="hbm", achieved=round(achieved, 1), peak=peak, unit="GB/s", frac=round(achieved / peak, 4),
                              traffic=(traffic or {}).get("dram_bytes_per_launch"), peak_source=peak_src,
                              kernel="search_kernel", kernel_ms=round(kern_avg_ms, 3),
                              algorithmic_bytes_per_launch=alg_bytes // args.steps),
                e2e=dict(value=round(e2e, 1), unit="queries/s", h2d_bytes_per_step=nq * wl["dim"] * 4,
                         d2h_bytes_per_step=nq * K * 8),
                gpu_launches=args.steps, clocks=clk, sweep=sweep)
    if block is not None:
        line["partitioned"] = block

    if world == 1:
        dumps = [d.tobytes() for d in ix.to_dumps(1)]
        qps, kind, done, secs, cpu_ids, cpu_dc = cpu_search(dumps, wl, batches[0].cpu().numpy(), ef, os.cpu_count(), args.cpu_seconds)
        line["cpu_baseline"] = dict(value=round(qps, 1), unit="queries/s", cores=os.cpu_count(), kind=kind, flags=CPU_ARM_FLAGS,
                                    sample=f"first {done} queries of batch 0 (ef={ef}, k={K}), {secs:.1f}s, {os.cpu_count()} threads x 4 coroutines")
        line["parity_at_scale"] = parity_at_scale(ix, batches[0], ef, cpu_ids, cpu_dc, stream)
        line["index_build_s"] = round(build_s, 1)
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()
    pas = line.get("parity_at_scale")
    if pas and pas["id_identical_frac"] < 0.99:
        raise SystemExit(f"parity at scale failed: {pas}")


if __name__ == "__main__":
    main()
