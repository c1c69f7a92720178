"""Scratch: partitioned search across 2 PROCESSES (CUDA IPC): rank 0 alone, then both ranks at once."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import torch.distributed as dist
import __graft_entry__ as ge
import bench
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
pkg = ge.load_package()
n, dim, nq = int(sys.argv[1]), int(sys.argv[3]) if len(sys.argv) > 3 else 128, 300000
base = bench.synth_rows(n, dim, 1001, dev)
full = pkg.Index.build_device(base.data_ptr(), n, dim, 16, 200, gpu=rank)
part = full.partition(rank, world, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
full.close(); del base
pkg.parallel.exchange_partition_shares(part, rank, world, dist, os.environ.get('MASTER_PORT', '0'))
dist.barrier()
import ctypes
L = pkg.shn.lib(); L.shn_debug_partition_gather_gbs.restype = ctypes.c_double; L.shn_debug_partition_gather_gbs.argtypes = [ctypes.c_void_p, ctypes.c_int]
if rank == 0:
    for prt in range(world):
        print(f'[rank 0] raw gather from share {prt}: {L.shn_debug_partition_gather_gbs(part._h, prt):.0f} GB/s', flush=True)
dist.barrier()
q = bench.synth_rows(nq, dim, 2002 + rank, dev)
ids = torch.empty((nq, 10), dtype=torch.int32, device=dev)
def run(tag):
    st = part.search_device(q.data_ptr(), nq, 10, 64, ids.data_ptr())
    tot = st["rows_hot"] + st["rows_local"] + st["rows_remote"]
    print(f"[rank {rank}] {tag}: {nq / st['kernel_ms'] / 1e3:.3f} MQPS, nvlink in {st['rows_remote'] * 4 * dim / st['kernel_ms'] / 1e6:.0f} GB/s", flush=True)
if rank == 0:
    run("alone (warm)"); run("alone")
dist.barrier()
run("both ranks at once"); run("both ranks at once")
dist.barrier()
dist.destroy_process_group()
