"""Test / bench helper (NOT part of the product path): the sharded cross-chain step over HOST buffers with any
torch.distributed backend.  The product's own multi-rank exchange is in C (ser_comm_* + ser_run_cross_chain_async:
NCCL on the run's stream; ser_multi_*: peer stores); this module exists so that the sharding rules -- global ids,
selection over the gathered E[-logL], disjoint pair-order slabs summed over ranks -- can be exercised with
world_size-2 `gloo` on a CPU-only box, and to broadcast the NCCL unique id under torchrun."""
import numpy as np


def cross_chain_distributed(S, e_local, k: int, po_counts_fn, n_sites: int, chains_selected=None, faithful: bool = True):
    """``e_local``: this rank's E[-logL] (global chain id = rank * len(e_local) + i);
    ``po_counts_fn(chosen_global_ids) -> int32 [k][N][N]`` fills the slabs of the chains this rank owns and leaves
    the others zero.  One all-gather of 8 bytes per chain, one all-reduce(sum) of the count slabs.
    Returns (chosen global ids, po matrix)."""
    import torch
    import torch.distributed as dist
    e_local = np.ascontiguousarray(e_local, dtype=np.float64)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if multi:
        world = dist.get_world_size()
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        mine = torch.from_numpy(e_local).to(dev)
        allv = torch.empty(world * e_local.size, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allv, mine)
        e_all = allv.cpu().numpy()
    else:
        dev, e_all = "cpu", e_local
    chosen, _, _ = S.select_chains(e_all, k)
    padded = np.full(k, -1, np.int32)
    padded[:len(chosen)] = chosen
    counts = np.ascontiguousarray(po_counts_fn(padded), dtype=np.int32)
    if multi:
        t = torch.from_numpy(counts).to(dev)
        dist.all_reduce(t)
        counts = t.cpu().numpy()
    po = S.po_finalize(counts[:max(1, len(chosen))], chains_selected or k, faithful) if len(chosen) else np.zeros((n_sites, n_sites))
    return [int(c) for c in chosen], po


def broadcast_comm_id(S, rank: int, device=None) -> bytes:
    """rank 0 creates the NCCL unique id of ser_comm_create; everyone receives its 128 bytes"""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(S.COMM_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        buf = torch.frombuffer(bytearray(S.Comm.unique_id()), dtype=torch.uint8).clone()
    if device is not None:
        buf = buf.to(device)
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())
