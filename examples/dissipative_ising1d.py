"""examples/dissipative_ising1d.jl of the reference through the nqcuda host mirror: steady state of the dissipative
Ising chain, NDM, sampled <L^dag L> + SR (Cholesky), observables from the diagonal chain.
usage: python examples/dissipative_ising1d.py [iterations]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neuralquantum.jl_b200"))
import nqcuda as nq  # noqa: E402

N, g, V = 7, 0.4, 2.0
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ctx = nq.Context(0)
hilb = nq.HomogeneousFock(N, 2)
H = nq.LocalOperator(hilb)
Sx, Sy, Sz = nq.LocalOperator(hilb), nq.LocalOperator(hilb), nq.LocalOperator(hilb)
ops = []
for i in range(1, N + 1):
    H = H + (g / 2.0) * nq.sigmax(hilb, i)
    H = H + ((V / 4.0) * nq.sigmaz(hilb, i)) * nq.sigmaz(hilb, i % N + 1)
    Sx = Sx + (1.0 / N) * nq.sigmax(hilb, i)
    Sy = Sy + (1.0 / N) * nq.sigmay(hilb, i)
    Sz = Sz + (1.0 / N) * nq.sigmaz(hilb, i)
    ops.append(nq.sigmam(hilb, i))
liouv = nq.liouvillian(H, ops)

sampl = nq.MetropolisSampler(nq.LocalRule(), 125, N, burn=100, seed=1234)
algo = nq.SR(np.float32, eps=0.001, algorithm=nq.sr_cholesky)
net = nq.NDM(ctx, hilb, np.float64, 1, 1, nq.af_softplus, seed=1234)
it = nq.BatchedSampler(net, sampl, liouv, algo, batch_sz=16)
for name, op in (("Sx", Sx), ("Sy", Sy), ("Sz", Sz)):
    it.add_observable_(name, op)
opt = nq.Descent(0.01)
for i in range(1, iters + 1):
    ldata, prec = it.sample_()
    ob = it.compute_observables()
    if i % 10 == 0 or i == 1:
        print("%d - %s   Sx %.4f  Sy %.4f  Sz %.4f" % (i, ldata, ob["Sx"].mean.real, ob["Sy"].mean.real, ob["Sz"].mean.real))
    it.precondition_(i)
    it.update_(opt)
