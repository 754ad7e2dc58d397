"""Oracle: stat_analysis (src/utils/stats.jl:26-50).  Test infrastructure only."""
import numpy as np


def stat_analysis(vals):
    """vals [B(chains), L].  Returns dict(mean, error, variance, tau, R).  Julia's var() is the
    corrected (n-1) estimator; var of complex = mean |x-mu|^2 corrected."""
    vals = np.asarray(vals)
    n, L = vals.shape

    def var(x, mean):
        x = np.asarray(x)
        return np.sum(np.abs(x - mean) ** 2) / (x.size - 1)
    mu_ch = vals.mean(axis=1)
    mu = mu_ch.mean()
    with np.errstate(invalid="ignore", divide="ignore"):
        var_ch = np.array([var(vals[i], mu_ch[i]) for i in range(n)])
        var_mu_ch = var(mu_ch, mu_ch.mean())
        var_mu = var(vals, mu)
        err = np.sqrt(var_mu_ch / n)
        t = var_mu_ch / var_mu
        tau = max(0.0, 0.5 * (t * L - 1))
        R = np.sqrt((L - 1) / L + t)
    return dict(mean=mu, error=err, variance=var_ch.mean(), tau=tau, R=R)
