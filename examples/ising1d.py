"""examples/ising1d.jl of the reference through the nqcuda host mirror: 1D transverse-field Ising ground state, RBM,
Metropolis local flips + SR (CG).  usage: python examples/ising1d.py [iterations]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neuralquantum.jl_b200"))
import nqcuda as nq  # noqa: E402

N, h, J = 20, 1.0, 1.0
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
ctx = nq.Context(0)
hilb = nq.HomogeneousSpin(N)
H = nq.LocalOperator(hilb)
Mx = nq.LocalOperator(hilb)
for i in range(1, N + 1):
    H = H - h * nq.sigmax(hilb, i)
    H = H + (J * nq.sigmaz(hilb, i)) * nq.sigmaz(hilb, i % N + 1)
    Mx = Mx + (1.0 / N) * nq.sigmax(hilb, i)

net = nq.RBM(ctx, hilb, np.float32, 1, nq.af_logcosh)
nq.init_random_pars_(net, sigma=0.01, seed=1234)
sampl = nq.MetropolisSampler(nq.LocalRule(), 125, N, burn=100, seed=1234)
algo = nq.SR(np.float32, eps=0.1, algorithm=nq.sr_cg, precision=1e-3)
it = nq.BatchedSampler(net, sampl, H, algo, batch_sz=8)
it.add_observable_("Mx", Mx)
opt = nq.Descent(0.1)
exact = -1.274549484318 * 20
for i in range(1, iters + 1):
    ldata, prec = it.sample_()
    ob = it.compute_observables()
    if i % 10 == 0 or i == 1:
        print("%d - %s   <Mx> = %.4f   (exact E = %.4f)" % (i, ldata, ob["Mx"].mean.real, exact))
    it.precondition_(i)
    it.update_(opt)
