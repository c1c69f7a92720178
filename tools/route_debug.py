"""Scratch (torchrun, N GPUs): routed step vs the unrouted whole-index answers, mismatches split by destination rank."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import torch.distributed as dist
import __graft_entry__ as ge
import bench

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
pkg = ge.load_package()
n, dim, nq, ef = int(sys.argv[1]), 128, int(sys.argv[2]), 64
use_visits = len(sys.argv) > 3 and sys.argv[3] == "visits"
K = 10
base = bench.synth_rows(n, dim, 1001, dev)
torch.cuda.synchronize()  # the library builds on its own stream
ix = pkg.Index.build_device(base.data_ptr(), n, dim, 16, 200, gpu=lr)
q = bench.synth_rows(nq, dim, 2002 + 1000 * rank, dev)
ref_i = torch.empty((nq, K), dtype=torch.int32, device=dev); ref_d = torch.empty((nq, K), dtype=torch.float32, device=dev)
ix.search_device(q.data_ptr(), nq, K, ef, ref_i.data_ptr(), ref_d.data_ptr())
# are the per-rank builds the same graph?  the same probe queries on every rank's own full index
probe = bench.synth_rows(20000, dim, 4242, dev)
pr_i = torch.empty((20000, K), dtype=torch.int32, device=dev)
torch.cuda.synchronize()
ix.search_device(probe.data_ptr(), 20000, K, ef, pr_i.data_ptr())
cs = (pr_i.long() * (torch.arange(20000 * K, device=dev).view(20000, K) % 1000003 + 1)).sum().view(1)
allcs = [torch.zeros_like(cs) for _ in range(world)]; dist.all_gather(allcs, cs)
print(f"[{rank}] probe result checksum per rank {[int(x) for x in allcs]} (equal = identical graphs)", flush=True)
if use_visits:
    ix.count_visits(True)
    torch.cuda.synchronize()
    warm = bench.synth_rows(100000, dim, 7007 + rank, dev)
    tmp = torch.empty((100000, K), dtype=torch.int32, device=dev)
    ix.search_device(warm.data_ptr(), 100000, K, ef, tmp.data_ptr())
    counts = torch.empty(ix.n, dtype=torch.int32, device=dev)
    ix.visit_counts(counts.data_ptr()); dist.all_reduce(counts); torch.cuda.synchronize(); ix.visit_counts(counts.data_ptr(), write_back=True)
owner = torch.empty(n, dtype=torch.uint8, device=dev)
cent, sizes = ix.placement_fit(world, owner.data_ptr(), seed=1234, slack=0.05)
oh = torch.zeros(1, dtype=torch.int64, device=dev) + int(owner.long().mul(torch.arange(n, device=dev) % 1000003).sum())
both = [torch.zeros_like(oh) for _ in range(world)]; dist.all_gather(both, oh)
print(f"[{rank}] owner checksum {[int(x) for x in both]} sizes {sizes.tolist()}", flush=True)
part = ix.partition(rank, world, 8, d_owner=owner.data_ptr())
ix.close()
pkg.parallel.exchange_partition_shares(part, rank, world, dist, "dbg" + os.environ.get("MASTER_PORT", "0"))
# 1) unrouted search on the partitioned handle
a_i = torch.empty((nq, K), dtype=torch.int32, device=dev)
st = part.search_device(q.data_ptr(), nq, K, ef, a_i.data_ptr())
print(f"[{rank}] unrouted on the partition: identical rows {(a_i == ref_i).all(1).float().mean().item():.4f} remote {st['rows_remote'] / max(1, st['rows_hot'] + st['rows_local'] + st['rows_remote']):.3f}", flush=True)
dist.barrier()
router = pkg.Router(part, cent, slack=0.25, max_batch=nq, k_max=K)
fd, size, _ = router.export(want_fd=True)
for peer, (pfd, psize) in bench.exchange_fds(fd, size, rank, world, dist, "dbgr" + os.environ.get("MASTER_PORT", "0")).items():
    router.attach(peer, fd=pfd, size=psize); os.close(pfd)
os.close(fd)
dist.barrier()
p_ids, p_d = router.results()
land = pkg.device_view(p_ids, (nq, K), "<i4")
for it in range(3):
    land.fill_(-7)
    torch.cuda.synchronize(); dist.barrier()
    router.scatter(q.data_ptr(), nq)
    torch.cuda.synchronize(); dist.barrier()
    st = router.search(K, ef)
    torch.cuda.synchronize(); dist.barrier()
    dest = pkg.device_view(router.destinations(), (nq,), "|u1").long()
    same = (land == ref_i).all(1)
    untouched = (land == -7).all(1)
    msg = [f"[{rank}] it{it} routed: identical {same.float().mean().item():.4f} untouched {untouched.float().mean().item():.4f} processed {st['processed']}"]
    for d in range(world):
        m = dest == d
        msg.append(f"dest{d}: n={int(m.sum())} identical {same[m].float().mean().item():.4f}")
    print("  ".join(msg), flush=True)
dist.barrier()
router.close(); part.close()
dist.destroy_process_group()
