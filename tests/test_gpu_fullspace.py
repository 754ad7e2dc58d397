"""Full-space tools on the device (csrc/nq_fullspace.cu) against the oracle restatement (oracle/fullspace.py):
ket / densitymatrix (utils/densitymatrix.jl:9-62) and the ExactSampler's table and draws (Samplers/Exact.jl:135-181)."""
import numpy as np
import pytest

import helpers as H
from oracle import fullspace as OF, machines as OM

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,hk,N,dt,act", [("rbm", "spin", 6, np.complex128, OM.LOGCOSH), ("rbm", "fock", 5, np.float64, OM.SOFTPLUS),
                                              ("rbm", "spin", 6, np.complex64, OM.LOGCOSH), ("rbm", "spin", 11, np.complex128, OM.LOGCOSH)])
def test_ket(nq, ctx, kind, hk, N, dt, act):
    om, pm, hilb = H.make_pair(nq, ctx, kind, hk, N, 2, dt, act, seed=11, std=0.3)
    oh = H.ohilb(hk, N)
    tol = H.TOL[np.dtype(dt)]
    for norm in (True, False):
        psi = nq.ket(pm, hilb, norm)
        H.assert_close(psi, OF.ket(om, oh, norm), tol, "ket norm=%s" % norm)
    assert abs(np.linalg.norm(nq.ket(pm, hilb)) - 1) < 10 * tol


@pytest.mark.parametrize("kind,N,dt", [("ndm", 3, np.float64), ("rbmsplit", 3, np.complex128), ("ndm", 3, np.float32), ("ndm", 5, np.float64)])
def test_densitymatrix(nq, ctx, kind, N, dt):
    om, pm, hilb = H.make_pair(nq, ctx, kind, "fock", N, 2, dt, OM.SOFTPLUS, seed=12, std=0.3)
    oh = H.ohilb("fock", N)
    tol = H.TOL[np.dtype(dt)]
    for norm in (True, False):
        rho = nq.densitymatrix(pm, hilb, norm)
        H.assert_close(rho, OF.densitymatrix(om, oh, norm), tol, "rho norm=%s" % norm)
    rho = nq.densitymatrix(pm, hilb)
    assert abs(np.trace(rho) - 1) < 10 * tol
    if kind == "ndm":                       # NDM is Hermitian and positive by construction (NDM.jl:76-96)
        assert np.abs(rho - rho.conj().T).max() < 10 * tol
        assert np.linalg.eigvalsh(rho.astype(np.complex128)).min() > -10 * tol
    v = nq.ket(pm, hilb)                    # ket(::MatrixNet) = normalised vec(rho)
    raw = nq.densitymatrix(pm, hilb, False)
    H.assert_close(v, raw.reshape(-1, order="F") / np.linalg.norm(raw), 10 * tol, "vec rho")


@pytest.mark.parametrize("kind,hk,N,dt", [("rbm", "spin", 10, np.complex128), ("ndm", "fock", 5, np.float64), ("rbmsplit", "fock", 4, np.complex128),
                                          ("rbm", "spin", 21, np.complex128)])
def test_exact_table_and_draws(nq, ctx, kind, hk, N, dt):
    """cdf to 1e-12 against the oracle's sequential cumsum; replayed draws: indices bit-equal to searchsortedfirst on the
    device table for every uniform, and equal to the oracle's for every uniform that is not within 1e-12 of a table entry."""
    import torch
    act = OM.LOGCOSH if kind == "rbm" else OM.SOFTPLUS
    om, pm, hilb = H.make_pair(nq, ctx, kind, hk, N, 1 if N > 12 else 2, dt, act, seed=13, std=0.3)
    big = N > 12
    cache = nq.ExactSamplerCache(nq.ExactSampler(64, seed=3), pm, 128)
    cdf = cache.init_sampler().cpu().numpy()
    assert cdf[-1] == 1.0 and np.all(np.diff(cdf) >= 0)
    rng = np.random.default_rng(5)
    Ls, B = 7, 128
    u = rng.random((Ls, B))
    u[0, :4] = [0.0, cdf[0], cdf[len(cdf) // 2], np.nextafter(1.0, 0)]         # edges: first entry, exact hits, last bin
    dev = torch.device("cuda", ctx.device)
    prow = torch.zeros((Ls, B, 1), dtype=torch.int64, device=dev)
    pcol = torch.zeros_like(prow) if pm.doubled else None
    idx = np.zeros((Ls, B), dtype=np.int64)
    cache.sample_into(Ls, prow, pcol, uniforms=u, indices=idx)
    assert np.array_equal(idx, OF.exact_draw(cdf, u))
    code = prow.cpu().numpy()[:, :, 0] + ((pcol.cpu().numpy()[:, :, 0] << N) if pm.doubled else 0)
    assert np.array_equal(code + 1, idx)
    if not big:
        ocdf = OF.exact_cdf(om, H.ohilb(hk, N))
        assert np.abs(cdf - ocdf).max() < 1e-12
        oidx = OF.exact_draw(ocdf, u)
        safe = np.abs(ocdf[np.minimum(oidx - 1, len(ocdf) - 1)] - u) > 1e-12
        safe &= (oidx < 2) | (np.abs(ocdf[np.maximum(oidx - 2, 0)] - u) > 1e-12)
        assert safe.mean() > 0.95 and np.array_equal(idx[safe], oidx[safe])
    # production draws: reproducible, independent of the sharding of the chains
    a = torch.zeros((Ls, B, 1), dtype=torch.int64, device=dev)
    ac = torch.zeros_like(a) if pm.doubled else None
    c1 = nq.ExactSamplerCache(nq.ExactSampler(64, seed=3), pm, B)
    c1.sample_into(Ls, a, ac)
    for off in (0, B // 2):
        b = torch.zeros((Ls, B // 2, 1), dtype=torch.int64, device=dev)
        bc = torch.zeros_like(b) if pm.doubled else None
        c2 = nq.ExactSamplerCache(nq.ExactSampler(64, seed=3), pm, B // 2, chain_offset=off)
        c2.sample_into(Ls, b, bc)
        assert torch.equal(b, a[:, off:off + B // 2])
