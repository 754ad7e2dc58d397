"""Run under torchrun with >= 2 ranks: the sharded iteration (NCCL all-reduce of <O>, F, S inside
libnqcuda) must reproduce the oracle's single-worker result on the union of the shards."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "neuralquantum.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import nqcuda as nq  # noqa: E402
import helpers as H  # noqa: E402
from oracle import machines as OM, sr as OSR, stats as OST  # noqa: E402
from oracle.models import lindblad_ising_1d  # noqa: E402

world, rank, local = nq.world_from_env()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = nq.Context(local, torch.cuda.current_stream().cuda_stream)
assert nq.init_comm(ctx) == (world, rank)
N, Btot, Lc = 6, 8 * world, 5
_, _, _, ol = lindblad_ising_1d(N)
_, _, _, pl = H.p_lindblad_ising_1d(nq, N)
om, pm, hilb = H.make_pair(nq, ctx, "ndm", "fock", N, 2, np.float64, OM.SOFTPLUS)
off, B = nq.shard_chains(Btot, world, rank)
R = H.rand_states("fock", N, Btot * Lc, 21).reshape(N, Btot, Lc, order="F")
Cc = H.rand_states("fock", N, Btot * Lc, 22).reshape(N, Btot, Lc, order="F")
eps = 0.001
for algo, tol in (("sr_cholesky", 1e-9), ("sr_cg", 1e-7)):
    bs = nq.BatchedSampler(pm, nq.MetropolisSampler(nq.LocalRule(), Lc * world, N, burn=5, seed=3), pl,
                           nq.SR(np.float32, eps=eps, algorithm=algo, precision=1e-12), batch_sz=B, chain_length=Lc)
    assert bs.Ns_total == Btot * Lc
    bs.set_samples((np.asfortranarray(R[:, off:off + B]), np.asfortranarray(Cc[:, off:off + B])))
    stat, _ = bs.sample_(sample=False)
    ref = OSR.iteration_liouvillian(om, ol, R.reshape(N, -1, order="F"), Cc.reshape(N, -1, order="F"), OSR.eps_f32(eps))
    if bs.S is not None:            # compare before the solve: Cholesky factorises S in place
        H.assert_close(bs.S.cpu().numpy().T, ref["S"], 1e-11, "S (global)")
    dw = bs.precondition_().cpu().numpy()
    H.assert_close(bs.avg.cpu().numpy(), ref["O_avg"], 1e-11, "<O> (global)")
    H.assert_close(bs.F.cpu().numpy(), ref["F"], 1e-9, "F (global)")
    assert abs(bs.cost - ref["C"]) < 1e-11 * ref["C"]
    # statistics are those of the union of the ranks' chains, identical on every rank (they were rank-local)
    gst = OST.stat_analysis((np.abs(ref["Lloc"]) ** 2).reshape(Btot, Lc, order="F"))
    for key in ("mean", "error", "variance", "tau", "R"):
        got, want = getattr(stat, key), gst[key]
        assert abs(got - want) <= 1e-10 * max(1.0, abs(want)), ("stat." + key, got, want)
    assert np.linalg.norm(dw - ref["dw"]) <= tol * np.linalg.norm(ref["dw"]), algo
    # every rank holds the same update
    t = torch.from_numpy(dw.copy()).cuda()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "ranks disagree on dw"
# uneven shards (chains do not divide by the number of ranks): offsets and the global sample count come from the
# all-reduced counts, every rank normalises <O>, F, S by the same number (overlapping chain ids, wrong S)
Bu = 8 * world + 1
offu, Bloc = nq.shard_chains(Bu, world, rank)
Ru = H.rand_states("fock", N, Bu * Lc, 31).reshape(N, Bu, Lc, order="F")
Cu = H.rand_states("fock", N, Bu * Lc, 32).reshape(N, Bu, Lc, order="F")
bs = nq.BatchedSampler(pm, nq.MetropolisSampler(nq.LocalRule(), Lc * world, N, burn=5, seed=3), pl,
                       nq.SR(np.float32, eps=eps, algorithm="sr_cholesky"), batch_sz=Bloc, chain_length=Lc)
assert (bs.chain_offset, bs.B_total, bs.Ns_total) == (offu, Bu, Bu * Lc), (bs.chain_offset, bs.B_total, offu, Bu)
bs.set_samples((np.asfortranarray(Ru[:, offu:offu + Bloc]), np.asfortranarray(Cu[:, offu:offu + Bloc])))
stat, _ = bs.sample_(sample=False)
ref = OSR.iteration_liouvillian(om, ol, Ru.reshape(N, -1, order="F"), Cu.reshape(N, -1, order="F"), OSR.eps_f32(eps))
H.assert_close(bs.S.cpu().numpy().T, ref["S"], 1e-11, "S (uneven shards)")
H.assert_close(bs.avg.cpu().numpy(), ref["O_avg"], 1e-11, "<O> (uneven shards)")
H.assert_close(bs.F.cpu().numpy(), ref["F"], 1e-9, "F (uneven shards)")
gst = OST.stat_analysis((np.abs(ref["Lloc"]) ** 2).reshape(Bu, Lc, order="F"))
assert abs(stat.mean - gst["mean"]) <= 1e-10 * abs(gst["mean"]) and abs(stat.error - gst["error"]) <= 1e-10 * gst["error"]
ctx.set_global_samples(0)
# sampler: the union of the shards equals one big run (Philox keyed by global chain id)
smp = nq.MetropolisSampler(nq.LocalRule(), 3, N, burn=4, seed=11)
part = nq.MetropolisSamplerCache(smp, pm, B, chain_offset=off)
part.randomize()
pr, pc = part.sample()
full = nq.MetropolisSamplerCache(smp, pm, Btot)
full.randomize()
fr, fc = full.sample()
assert np.array_equal(pr, fr[:, off:off + B]) and np.array_equal(pc, fc[:, off:off + B])
dist.barrier()
if rank == 0:
    print("MULTI_GPU_OK world=%d" % world)
dist.destroy_process_group()
