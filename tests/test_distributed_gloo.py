"""world_size-2 gloo test of the sharded cross-chain step (host logic only, no GPU): the chosen
chains and the pair-order matrix must equal the single-process result, whatever the sharding."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_inputs():
    rng = np.random.default_rng(11)
    n_chains, N, T, k = 16, 13, 25, 3
    e = rng.normal(4000, 30, n_chains)
    pis = np.array([[rng.permutation(N) for _ in range(T)] for _ in range(n_chains)])
    return e, pis, N, T, k


def _counts_for(pis, owned, chosen):
    from oracle import oracle as O
    out = np.zeros((len(chosen), pis.shape[2], pis.shape[2]), np.int32)
    for slot, g in enumerate(chosen):
        if g >= 0 and g in owned:
            out[slot] = O.pair_order_counts(pis[g])
    return out


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import seriation_b200 as S
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    e, pis, N, T, k = _make_inputs()
    per = len(e) // world
    owned = set(range(rank * per, (rank + 1) * per))
    from tools.dist_helpers import cross_chain_distributed
    chosen, po = cross_chain_distributed(S, e[rank * per:(rank + 1) * per], k, lambda ch: _counts_for(pis, owned, ch), N)
    q.put((rank, chosen, po))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_selection_and_pair_order_equal_single_process():
    sys.path.insert(0, ROOT)
    import seriation_b200 as S
    from oracle import oracle as O
    e, pis, N, T, k = _make_inputs()
    want_chosen = O.choose_chains(e, k)
    want_po = O.pair_order_matrix([O.pair_order_counts(pis[g]) for g in want_chosen], k)
    from tools.dist_helpers import cross_chain_distributed
    single_chosen, single_po = cross_chain_distributed(S, e, k, lambda ch: _counts_for(pis, set(range(len(e))), ch), N)
    assert single_chosen == want_chosen and np.allclose(single_po, want_po, atol=1e-15)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, chosen, po in results:
        assert chosen == want_chosen, rank
        assert np.allclose(po, want_po, atol=1e-15), rank
