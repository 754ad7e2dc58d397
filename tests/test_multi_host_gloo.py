"""world_size-2 CPU (gloo) tests of the multi-GPU recipe's HOST logic (prompt section 5):
chain sharding, and the reduction recipe libnqcuda uses under sharding -- every rank normalises its
partial sums by the GLOBAL sample count and the partials are all-reduced with SUM (C1-C4, quirk Q5) --
checked with the oracle standing in for the per-rank device work."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "neuralquantum.jl_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nqcuda.parallel import shard_chains
    from oracle import machines as OM, sr as OSR, estimators as OE
    from oracle.models import lindblad_ising_1d, random_states
    N, B, Lc = 4, 6, 5
    hilb, _, _, liouv = lindblad_ising_1d(N)
    net = OM.random_machine("ndm", N, 1, seed=3, std=0.3)
    R = random_states(hilb, B * Lc, 1).reshape(N, B, Lc, order="F")
    Cc = random_states(hilb, B * Lc, 2).reshape(N, B, Lc, order="F")
    off, cnt = shard_chains(B, world, rank)
    r = R[:, off:off + cnt].reshape(N, -1, order="F")
    c = Cc[:, off:off + cnt].reshape(N, -1, order="F")
    Ns_tot = B * Lc
    out, O = net.logpsi_grad(r, c)
    Lloc, gL = OE.local_grad_super(net, liouv, r, c, out)

    def allsum(x):
        t = torch.from_numpy(np.ascontiguousarray(x).view(np.float64).copy())
        dist.all_reduce(t)
        return t.numpy().view(x.dtype).reshape(x.shape)
    avg = allsum(O.sum(1)) / Ns_tot                                   # C1: global <O>
    Oc = O - avg[:, None]
    packed = np.concatenate([(gL.conj() * Lloc[None, :]).sum(1), [np.sum(np.abs(Lloc) ** 2) + 0j]]) / Ns_tot
    packed = allsum(packed)                                           # C2 + C3 packed
    cost = packed[-1].real
    F = (packed[:-1].conj() - cost * avg).real
    S = allsum((np.real(Oc @ Oc.conj().T) / Ns_tot))                  # C4: partial S normalised by global Ns, SUM
    if rank == 0:
        ref = OSR.iteration_liouvillian(net, liouv, R.reshape(N, -1, order="F"), Cc.reshape(N, -1, order="F"), 1e-3)
        q.put((float(np.abs(S - ref["S"]).max()), float(np.abs(F - ref["F"]).max()), float(abs(cost - ref["C"])),
               float(np.abs(avg - ref["O_avg"]).max())))
    dist.destroy_process_group()


def test_sharded_reduction_recipe_matches_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert max(res) < 1e-12, res


def test_shard_chains_partition():
    sys.path.insert(0, os.path.join(ROOT, "neuralquantum.jl_b200"))
    from nqcuda.parallel import shard_chains
    for total in (1, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            parts = [shard_chains(total, world, r) for r in range(world)]
            assert sum(c for _, c in parts) == total
            assert parts[0][0] == 0
            for (o1, c1), (o2, _) in zip(parts, parts[1:]):
                assert o1 + c1 == o2
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        shard_chains(8, 2, 2)
