"""GPU parity at BASELINE's full size (cfg4: dissipative Ising N=16, NDM alpha=2, 65 536 configurations per
iteration), where the NumPy oracle would take hours: size-independent properties + the oracle on a random subset.

  * two independently written kernels (site-local fused kernel vs ordered-connection-list kernel) agree on every
    configuration; a random subset of the batch is checked against the oracle;
  * batch-permutation equivariance (bitwise);
  * S against an independent FP64 matmul of the centred rows (checksum S v, diagonal, symmetry), FP64 and FP32 mode;
  * the sampler does not depend on how the chains are sharded (bitwise)."""
import os

import numpy as np
import pytest
import torch

import helpers as H
from oracle import machines as OM
from oracle import estimators as OE
from oracle.models import lindblad_ising_1d

pytestmark = pytest.mark.gpu

N, B, LC = 16, 4096, 16


def _setup(nq, ctx, dtype, seed=7):
    _, _, _, pl = H.p_lindblad_ising_1d(nq, N)
    om, pm, hilb = H.make_pair(nq, ctx, "ndm", "fock", N, 2, dtype, OM.SOFTPLUS)
    bs = nq.BatchedSampler(pm, nq.MetropolisSampler(nq.LocalRule(), LC, N, burn=4, seed=seed), pl,
                           nq.SR(np.float32, eps=1e-3, algorithm="sr_cholesky"), batch_sz=B)
    rng = np.random.default_rng(seed)
    sig = (rng.integers(0, 2, (N, B, LC)).astype(float), rng.integers(0, 2, (N, B, LC)).astype(float))
    bs.set_samples(sig)
    return om, pm, bs, sig


def test_two_kernels_agree_and_match_oracle_subset(nq, ctx):
    om, pm, bs, sig = _setup(nq, ctx, np.float64)
    bs.evaluate()
    torch.cuda.synchronize()
    a = [t.clone() for t in (bs.logpsi, bs.O, bs.loc, bs.gloc)]
    os.environ["NQ_NDM3_KERNEL"] = "list"
    try:
        bs.evaluate()
        torch.cuda.synchronize()
    finally:
        del os.environ["NQ_NDM3_KERNEL"]
    b = (bs.logpsi, bs.O, bs.loc, bs.gloc)
    for x, y, name in zip(a, b, ("log rho", "O", "L_loc", "grad L_loc")):
        scale = float(y.abs().max())
        assert float((x - y).abs().max()) <= 1e-11 * max(scale, 1.0), name
    # oracle on 96 random configurations of the batch
    _, _, _, ol = lindblad_ising_1d(N)
    rng = np.random.default_rng(3)
    idx = rng.choice(B * LC, 96, replace=False)
    # configuration s of the flat batch = (chain b, slot l) with s = l * B + b  (set_samples layout [N, B, L])
    R = sig[0].reshape(N, B * LC, order="F")[:, idx]
    Cc = sig[1].reshape(N, B * LC, order="F")[:, idx]
    ref_lp = om.logpsi(R, Cc)
    ref_loc, ref_g = OE.local_grad_super(om, ol, R, Cc)
    tidx = torch.from_numpy(idx).cuda()
    H.assert_close(a[0][tidx].cpu().numpy(), ref_lp, 1e-11, "log rho (subset)")
    H.assert_close(a[2][tidx].cpu().numpy(), ref_loc, 1e-11, "L_loc (subset)")
    H.assert_close(a[3][tidx].cpu().numpy().T, ref_g, 1e-11, "grad L_loc (subset)")


def test_batch_permutation_equivariance(nq, ctx):
    om, pm, bs, sig = _setup(nq, ctx, np.float64)
    bs.evaluate()
    torch.cuda.synchronize()
    loc0, g0, o0 = bs.loc.clone(), bs.gloc.clone(), bs.O.clone()
    rng = np.random.default_rng(5)
    perm = rng.permutation(B)
    bs.set_samples((sig[0][:, perm, :], sig[1][:, perm, :]))
    bs.evaluate()
    torch.cuda.synchronize()
    # flat index s = l * B + b
    tperm = torch.from_numpy(np.concatenate([l * B + perm for l in range(LC)])).cuda()
    assert torch.equal(bs.loc, loc0[tperm])
    assert torch.equal(bs.gloc, g0[tperm])
    assert torch.equal(bs.O, o0[tperm])


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-11), (np.float32, 1e-5)])
def test_s_matrix_against_independent_matmul(nq, ctx, dtype, tol):
    om, pm, bs, sig = _setup(nq, ctx, dtype)
    bs.sample_(sample=False)              # evaluate + centre + force + S
    torch.cuda.synchronize()
    P, Ns = pm.P, B * LC
    Oc = bs.O.to(torch.complex128)        # [Ns, P] = centred rows, column-major [P, Ns]
    S = bs.S.to(torch.float64).clone()          # the Cholesky solve factors bs.S in place
    assert torch.equal(bs.S, bs.S.T)
    assert float(Oc.sum(dim=0).abs().max()) <= (1e-9 if dtype == np.float64 else 1e-1)      # centred
    g = torch.Generator(device="cuda").manual_seed(1)
    v = torch.randn(P, 4, dtype=torch.float64, device="cuda", generator=g)
    # S = Re(Oc^T conj(Oc)) / Ns in this layout:  S v = Re(Oc^T (conj(Oc) v)) / Ns
    ref = (Oc.T @ (Oc.conj() @ v.to(torch.complex128))).real / Ns
    got = S @ v
    assert float((got - ref).abs().max()) <= tol * float(ref.abs().max()), "S v checksum"
    dref = (Oc.abs() ** 2).sum(dim=0) / Ns
    assert float((torch.diagonal(S) - dref).abs().max()) <= tol * float(dref.max()), "diag S"
    if dtype == np.float64:
        dw = bs.precondition_()            # Cholesky succeeds: S + eps I is positive definite
        assert torch.isfinite(dw).all()
        dw = dw.to(torch.float64)
        r = S @ dw + float(np.float32(1e-3)) * dw - bs.F.to(torch.float64)
        scale = float(S.abs().sum(dim=1).max()) * float(dw.abs().max()) + float(bs.F.abs().max())
        assert float(r.abs().max()) <= 1e-11 * scale


def test_sampler_independent_of_sharding(nq, ctx):
    from nqcuda.samplers import MetropolisSamplerCache
    om, pm, hilb = H.make_pair(nq, ctx, "ndm", "fock", N, 2, np.float64, OM.SOFTPLUS)
    smp = nq.MetropolisSampler(nq.LocalRule(), 4, N, burn=6, seed=21)
    W = nq._lib.lib.nq_states_words(N)

    def run(nchains, offset):
        cache = MetropolisSamplerCache(smp, pm, nchains, chain_offset=offset)
        prow = torch.zeros((4, nchains, W), dtype=torch.int64, device="cuda")
        pcol = torch.zeros_like(prow)
        cache.randomize()
        cache.sample(smp.burn_length, 4, packed_out=(prow.data_ptr(), pcol.data_ptr()))
        torch.cuda.synchronize()
        return prow, pcol
    rows, cols = run(B, 0)
    half = B // 2
    for r in range(2):
        pr, pc = run(half, r * half)
        assert torch.equal(pr, rows[:, r * half:(r + 1) * half])
        assert torch.equal(pc, cols[:, r * half:(r + 1) * half])
    assert int((rows != 0).sum()) > 0


def test_pipelined_host_evaluate_is_bit_identical(nq, ctx):
    """BatchedSampler.evaluate_host (copies and packing of piece c+1 on a side stream while the fused kernel works on piece c)
    gives exactly the arrays of set_samples + evaluate."""
    import torch
    N, B, Lc = 8, 4096, 10                         # 40 960 configurations = 17.3 rounds of 148 x 16 warps: three pieces by default
    om, pm, hilb = H.make_pair(nq, ctx, "ndm", "fock", N, 2, np.float64, 0, seed=4, std=0.2)
    _, _, _, pl = H.p_lindblad_ising_1d(nq, N)
    bs = nq.BatchedSampler(pm, nq.MetropolisSampler(nq.LocalRule(), Lc, N, burn=1, seed=1), pl, nq.SR(np.float32, eps=0.001), batch_sz=B)
    Ns = B * Lc
    hr = torch.from_numpy(H.rand_states("fock", N, Ns, 11).T.copy()).pin_memory()
    hc = torch.from_numpy(H.rand_states("fock", N, Ns, 12).T.copy()).pin_memory()
    sig = (hr.numpy().T.reshape(N, B, Lc, order="F"), hc.numpy().T.reshape(N, B, Lc, order="F"))
    bs.set_samples(sig)
    bs.evaluate()
    torch.cuda.synchronize()
    ref = [t.clone() for t in (bs.prow, bs.pcol, bs.logpsi, bs.O, bs.loc, bs.gloc)]
    for t in (bs.prow, bs.pcol, bs.logpsi, bs.O, bs.loc, bs.gloc):
        t.zero_()
    # the library entry takes HOST configurations only and says so (device arrays go through nq_pack_states + the packed entry)
    L = nq._lib
    with pytest.raises(nq.NQError):
        L.check(L.lib.nq_logpsi_grad_local_host(pm.h, bs.op.h, bs.prow.data_ptr(), bs.pcol.data_ptr(), L.NQ_F64, Ns, bs.prow.data_ptr(),
                                                bs.pcol.data_ptr(), bs.logpsi.data_ptr(), bs.O.data_ptr(), pm.P, bs.loc.data_ptr(),
                                                bs.gloc.data_ptr(), pm.P), ctx.h)
    for chunks in (2, None):                       # equal pieces from the mirror; the library's growing schedule (2 rounds, 7, rest)
        for t in (bs.prow, bs.pcol, bs.logpsi, bs.O, bs.loc, bs.gloc):
            t.zero_()
        bs.evaluate_host(sig, chunks=chunks)
        torch.cuda.synchronize()
        for a, b in zip(ref, (bs.prow, bs.pcol, bs.logpsi, bs.O, bs.loc, bs.gloc)):
            assert torch.equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.complex128, np.float32])
def test_host_entry_ket_machine_is_bit_identical(nq, ctx, dtype):
    """nq_logpsi_grad_local_host on a ket machine (RBM + Hamiltonian: no sigma', no gradient of the estimator; the library
    runs its two kernels per piece) against set_samples + evaluate, float32 and float64 host arrays."""
    import torch
    N, B, Lc = 12, 2048, 8                         # 16 384 configurations: two pieces (2 rounds + the rest)
    om, pm, hilb = H.make_pair(nq, ctx, "rbm", "spin", N, 2, dtype, OM.LOGCOSH, seed=7)
    _, pH = H.p_tfim_1d(nq, N)
    bs = nq.BatchedSampler(pm, nq.MetropolisSampler(nq.LocalRule(), Lc, N, burn=1, seed=1), pH, nq.SR(np.float32, eps=0.001), batch_sz=B)
    Ns = B * Lc
    for fdt in (np.float64, np.float32):
        sig = np.asfortranarray(H.rand_states("spin", N, Ns, 21).astype(fdt)).reshape(N, B, Lc, order="F")
        bs.set_samples(sig)
        bs.evaluate()
        torch.cuda.synchronize()
        ref = [t.clone() for t in (bs.prow, bs.logpsi, bs.O, bs.loc)]
        for t in (bs.prow, bs.logpsi, bs.O, bs.loc):
            t.zero_()
        bs.evaluate_host(sig)
        torch.cuda.synchronize()
        for a, b in zip(ref, (bs.prow, bs.logpsi, bs.O, bs.loc)):
            assert torch.equal(a, b)
        assert int((bs.loc != 0).sum()) > 0
