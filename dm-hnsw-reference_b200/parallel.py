"""Host-side logic of the multi-GPU path: one process per GPU (torch.distributed), queries sharded the way the
reference deals them to compute nodes (round-robin by id, src/io/read_data.hh:58), per-GPU top-k lists exchanged with
ONE all-gather per batch, recall combined as the reference's rolling recall (compute_node.cc:149-154)."""
import numpy as np
import torch


def shard_slots(n, rank, world):
    """Query ids this rank processes: id % world == rank (io/read_data.hh:58)."""
    return np.arange(rank, n, world, dtype=np.int64)


def shard_size(n, rank, world):
    return (n - rank + world - 1) // world if rank < n else 0


def allgather_results(ids, dists, n_total, rank, world, dist=None):
    """ids [n_local,k] int32, dists [n_local,k] float32 for the slots shard_slots(n_total, rank, world), on any device.
    Returns ([n_total,k] ids, [n_total,k] dists) in global query order on every rank.  One all-gather (ids and the
    distance bits travel in one int32 buffer)."""
    k = ids.shape[1]
    if world == 1:
        return ids, dists
    per = (n_total + world - 1) // world  # ranks with one slot fewer pad the last row
    send = torch.full((per, 2 * k), -1, dtype=torch.int32, device=ids.device)
    send[: ids.shape[0], :k] = ids
    send[: ids.shape[0], k:] = dists.view(torch.int32)
    recv = torch.empty((world, per, 2 * k), dtype=torch.int32, device=ids.device)
    dist.all_gather_into_tensor(recv.view(world * per, 2 * k), send)
    # slot s of rank r is global query r + s*world: [world, per] -> [per, world] -> flatten, cut the padding
    merged = recv.permute(1, 0, 2).reshape(per * world, 2 * k)[:n_total]
    return merged[:, :k].contiguous(), merged[:, k:].contiguous().view(torch.float32)


def local_recall(ids, gt_rows):
    """compute_local_recall (compute_node.cc:579-600): hits / (processed * k) against the first k ground-truth ids."""
    k = ids.shape[1]
    if ids.shape[0] == 0:
        return 0.0
    hit = (ids.long()[:, :, None] == gt_rows[:, None, :k].long()).any(2).sum().item()
    return hit / (ids.shape[0] * k)


def rolling_recall(recall_local, processed_local, n_total, dist=None, device="cpu"):
    """sum_i recall_i * processed_i / total (compute_node.cc:149-154, statistics.hh combine())."""
    t = torch.tensor([recall_local * processed_local / max(n_total, 1)], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t)
    return float(t.item())


def exchange_partition_shares(part, rank, world, dist, tag):
    """Every rank maps every other rank's share: the POSIX fds of the CUDA VMM allocations travel over Unix sockets
    (SCM_RIGHTS), one listening socket per rank in the abstract namespace."""
    import os
    import socket
    import threading
    fds, sizes, _ = part.partition_export(want_fds=True)
    name = lambda r: f"\0shn-share-{tag}-{r}"
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(name(rank))
    srv.listen(world)

    def serve():
        for _ in range(world - 1):
            conn, _ = srv.accept()
            with conn:
                msg = f"{sizes[0]} {sizes[1]}".encode()
                socket.send_fds(conn, [msg], fds)

    th = threading.Thread(target=serve, daemon=True)
    th.start()
    dist.barrier()  # every rank is listening
    for peer in range(world):
        if peer == rank:
            continue
        with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
            c.connect(name(peer))
            msg, got, _, _ = socket.recv_fds(c, 256, 2)
        psz = [int(x) for x in msg.decode().split()]
        part.partition_attach(peer, fds=got, sizes=psz)
        for fd in got:
            os.close(fd)  # the import holds its own reference
    th.join()
    srv.close()
    for fd in fds:
        os.close(fd)
    dist.barrier()


class RoutedExchange:
    """Query routing between ranks (the reference: QueryRouter::route_query / poll_recv_cq relayed through memory nodes,
    src/router/query_router.hh:83-104,195-210): rows travel to the rank `dest` names with ONE all-to-all(v), results come
    back with a second one and are put back in the caller's order.  Works on CPU (gloo) and GPU (nccl) tensors."""

    def __init__(self, dest, world, dist, device):
        self.world, self.dist = world, dist
        dest = torch.as_tensor(dest, dtype=torch.int64, device=device)
        self.order = torch.argsort(dest, stable=True)
        self.send_counts = torch.bincount(dest, minlength=world)
        recv = torch.empty_like(self.send_counts)
        dist.all_to_all_single(recv, self.send_counts)
        self.recv_counts = recv
        self.send_list = self.send_counts.tolist()
        self.recv_list = self.recv_counts.tolist()

    def forward(self, rows):
        """rows [n, ...] in the caller's order -> the rows this rank has to process."""
        out = torch.empty((sum(self.recv_list),) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
        self.dist.all_to_all_single(out, rows[self.order].contiguous(), self.recv_list, self.send_list)
        return out

    def backward(self, results):
        """results [n_received, ...] -> results for the caller's own rows, in the caller's order."""
        back = torch.empty((sum(self.send_list),) + tuple(results.shape[1:]), dtype=results.dtype, device=results.device)
        self.dist.all_to_all_single(back, results.contiguous(), self.send_list, self.recv_list)
        out = torch.empty_like(back)
        out[self.order] = back
        return out
