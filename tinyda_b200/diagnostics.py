"""Post-processing: ``get_samples`` / ``to_xarray`` / ``to_inference_data`` with the reference's
signatures and naming (tinyDA/diagnostics.py:6-209), reading the array-backed LinkSequence
directly instead of one Python object per sample, plus rank-normalised split-R-hat and bulk
ESS (Vehtari et al. 2021, what ArviZ computes on the reference's output) in NumPy so that
the benchmark's min-ESS/sec needs neither ArviZ nor the reference.
"""
import numpy as np

from .link import LinkSequence


def _attr_array(seq, attribute):
    if isinstance(seq, LinkSequence):
        if attribute == "parameters":
            return np.asarray(seq.parameters)
        if attribute == "model_output":
            if seq.model_output is None:
                raise ValueError("model outputs were not stored (store_model_output=False)")
            return np.asarray(seq.model_output)
        if attribute == "stats":
            return np.stack([seq.prior, seq.likelihood, seq.prior + seq.likelihood], axis=1)
        if attribute == "qoi":
            # link.py:38-48: None unless the model returns (output, qoi), posterior.py:95-105
            return np.array([None] * len(seq)) if seq.qoi is None else np.asarray(seq.qoi)
    if attribute == "stats":
        return np.array([[l.prior, l.likelihood, l.posterior] for l in seq])
    return np.array([getattr(l, attribute) for l in seq])


def get_samples(chain, attribute="parameters", level="fine", burnin=0):
    """diagnostics.py:114-209: dict with 'chain_i' arrays (samples as rows), 'iterations',
    'dimension' plus the sampler info copied across."""
    samples = {"sampler": chain["sampler"], "n_chains": chain["n_chains"], "attribute": attribute}
    if chain["sampler"] == "MH":
        key = "chain_{}"
    elif chain["sampler"] == "DA":
        samples["subchain_length"] = chain["subchain_length"]
        samples["level"] = level
        key = "chain_" + str(level) + "_{}"
    elif chain["sampler"] == "MLDA":
        samples["subchain_lengths"] = chain["subchain_lengths"]
        samples["level"] = level
        key = "chain_l" + str(level) + "_{}"
    else:
        raise ValueError("unknown sampler %r" % chain["sampler"])
    # under torch.distributed every rank holds its own shard of the chains (sample() records the range)
    lo, hi = chain.get("local_chains", (0, chain["n_chains"]))
    fast = _dense_from_history(chain, key, attribute, burnin)
    for i in range(lo, hi):
        if fast is not None:
            x = fast[i - lo]
        else:
            x = _attr_array(chain[key.format(i)][burnin:], attribute)
        if x.ndim == 1:
            x = x[..., np.newaxis]
        samples["chain_{}".format(i)] = x
    samples["iterations"] = samples["chain_{}".format(lo)].shape[0]
    samples["dimension"] = samples["chain_{}".format(lo)].shape[1]
    if (lo, hi) != (0, chain["n_chains"]):
        samples["local_chains"] = (lo, hi)
    return samples


def _dense_from_history(chain, key, attribute, burnin):
    """All chains of the finest level at once from the compacted history sample() attaches to its result
    (no per-chain Python objects); None when the request is for another level or a plain dict."""
    hist = getattr(chain, "history", None)
    if hist is None:
        return None
    fine_key = {"MH": "chain_{}", "DA": "chain_fine_{}"}.get(chain["sampler"], "chain_l%d_{}" % (chain.get("levels", 1) - 1))
    if key != fine_key:
        return None
    if attribute == "parameters":
        return hist.dense("theta", burnin)
    if attribute == "model_output":
        if hist.chunks and hist.chunks[0].output is None:
            raise ValueError("model outputs were not stored (store_model_output=False)")
        return hist.dense("output", burnin)
    if attribute == "stats":
        pr, lk = hist.dense("prior", burnin), hist.dense("like", burnin)
        return np.stack([pr, lk, pr + lk], axis=2)
    if attribute == "qoi" and hist.chunks and hist.chunks[0].qoi is not None:
        return hist.dense("qoi", burnin)
    return None


def to_xarray(samples, keys):
    """diagnostics.py:72-111 (needs xarray)."""
    import xarray as xr
    data_vars = {}
    for i in range(samples["dimension"]):
        lo, hi = samples.get("local_chains", (0, samples["n_chains"]))
        x = np.array([samples["chain_{}".format(j)][:, i] for j in range(lo, hi)])
        data_vars[keys[i]] = (["chain", "draw"], x)
    lo, hi = samples.get("local_chains", (0, samples["n_chains"]))
    return xr.Dataset(
        data_vars=data_vars,
        coords=dict(chain=("chain", list(range(lo, hi))),
                    draw=("draw", list(range(samples["iterations"])))),
    )


def to_inference_data(chain, level="fine", burnin=0, parameter_names=None):
    """diagnostics.py:6-69 (needs arviz + xarray): groups posterior, posterior_predictive, qoi,
    sample_stats; variables x{i}, obs_{i}, qoi_{i}, prior / likelihood / posterior."""
    import arviz as az
    arrays = []
    for attr in ["parameters", "model_output", "qoi", "stats"]:
        samples = get_samples(chain, attr, level, burnin)
        if attr == "parameters":
            keys = (["x{}".format(i) for i in range(samples["dimension"])]
                    if parameter_names is None else parameter_names)
        elif attr == "model_output":
            keys = ["obs_{}".format(i) for i in range(samples["dimension"])]
        elif attr == "qoi":
            keys = ["qoi_{}".format(i) for i in range(samples["dimension"])]
        else:
            keys = ["prior", "likelihood", "posterior"]
        arrays.append(to_xarray(samples, keys))
    return az.InferenceData(posterior=arrays[0], posterior_predictive=arrays[1], qoi=arrays[2],
                            sample_stats=arrays[3])


# ---- R-hat / ESS ---------------------------------------------------------------------------
def _split(x):
    n = x.shape[1] // 2
    return np.concatenate([x[:, :n], x[:, x.shape[1] - n:]], axis=0)


def _rank_normalise(x):
    from scipy.stats import rankdata, norm
    r = rankdata(x.reshape(-1), method="average").reshape(x.shape)
    return norm.ppf((r - 0.375) / (x.size + 0.25))


def _autocov(x):
    n = x.shape[1]
    m = 1 << int(np.ceil(np.log2(2 * n)))
    xc = x - x.mean(axis=1, keepdims=True)
    f = np.fft.rfft(xc, n=m, axis=1)
    ac = np.fft.irfft(f * np.conj(f), n=m, axis=1)[:, :n]
    return ac / n


def _rhat_plain(x):
    m, n = x.shape
    W = x.var(axis=1, ddof=1).mean()
    B = n * x.mean(axis=1).var(ddof=1)
    return np.sqrt(((n - 1) / n * W + B / n) / W)


def _ess_plain(x):
    m, n = x.shape
    acov = _autocov(x)
    means = x.mean(axis=1)
    return _ess_from_sums(m, n, acov.sum(axis=0), means.sum(), (means ** 2).sum())


def _ess_from_sums(m, n, acov_sum, mean_sum, mean_sq_sum):
    """Multi-chain ESS (Geyer's initial sequences, as in Stan / ArviZ) from quantities that are
    plain sums over chains -- so they can be all-reduced across GPUs: acov_sum[t] = sum over chains
    of the biased autocovariance at lag t, and the sums of the chain means and of their squares."""
    acov_mean = acov_sum / m
    mean_var = acov_mean[0] * n / (n - 1.0)
    var_plus = mean_var * (n - 1.0) / n
    if m > 1:
        var_plus += (mean_sq_sum - mean_sum ** 2 / m) / (m - 1.0)
    acov = acov_mean[None, :]
    rho = np.zeros(n)
    rho[0] = 1.0
    t = 1
    rho_even, rho_odd = 1.0, 1.0 - (mean_var - acov[:, 1].mean()) / var_plus
    rho[1] = rho_odd
    while t < n - 3 and (rho_even + rho_odd) > 0:       # Geyer's initial positive sequence
        rho_even = 1.0 - (mean_var - acov[:, t + 1].mean()) / var_plus
        rho_odd = 1.0 - (mean_var - acov[:, t + 2].mean()) / var_plus
        if rho_even + rho_odd >= 0:
            rho[t + 1], rho[t + 2] = rho_even, rho_odd
        t += 2
    max_t = t - 2
    if rho_even > 0:
        rho[max_t + 1] = rho_even
    t = 1
    while t <= max_t - 2:                                # initial monotone sequence
        if rho[t + 1] + rho[t + 2] > rho[t - 1] + rho[t]:
            rho[t + 1] = (rho[t - 1] + rho[t]) / 2.0
            rho[t + 2] = rho[t + 1]
        t += 2
    tau = -1.0 + 2.0 * rho[:max_t + 1].sum() + rho[max_t + 1]
    tau = max(tau, 1.0 / np.log10(m * n))
    return m * n / tau


def rhat(x):
    """Rank-normalised split-R-hat of draws x [n_chains, n_draws]."""
    x = np.asarray(x, dtype=np.float64)
    z = _rank_normalise(_split(x))
    zf = _rank_normalise(np.abs(_split(x) - np.median(x)))
    return max(_rhat_plain(z), _rhat_plain(zf))


def ess_bulk(x):
    """Bulk effective sample size of draws x [n_chains, n_draws] (rank-normalised, split)."""
    x = np.asarray(x, dtype=np.float64)
    return _ess_plain(_rank_normalise(_split(x)))


def ess_rhat_from_sums(sums, folded):
    """Bulk ESS and rank-normalised split R-hat per parameter from the device-side sums of
    ``Engine.ess_sums`` (tda_ess_sums): sums [d, n_lag + 4] = autocovariance sums over the split chains,
    then (sum of means, sum of squared means, split chains, n_half); folded [d, 4] = (lag-0 sum, sum of
    means, sum of squared means, split chains) of the folded scores.  Same estimators as ``ess_bulk`` /
    ``rhat`` above, which work on host arrays."""
    sums = np.atleast_2d(np.asarray(sums, dtype=np.float64))
    folded = np.atleast_2d(np.asarray(folded, dtype=np.float64))
    d, n_lag = sums.shape[0], sums.shape[1] - 4
    ess, rh = np.zeros(d), np.zeros(d)

    def _rhat(ac0_sum, mean_sum, mean_sq_sum, m, n):
        W = ac0_sum / m * n / (n - 1.0)
        B_over_n = (mean_sq_sum - mean_sum ** 2 / m) / (m - 1.0)
        return np.sqrt(((n - 1.0) / n * W + B_over_n) / W)

    for k in range(d):
        m, n = int(round(sums[k, n_lag + 2])), int(round(sums[k, n_lag + 3]))
        ac = np.zeros(n)
        ac[:n_lag] = sums[k, :n_lag]                   # lags beyond n_lag count as zero correlation mass
        ess[k] = _ess_from_sums(m, n, ac, sums[k, n_lag], sums[k, n_lag + 1]) if n_lag == n else \
            _ess_from_sums_truncated(m, n, sums[k, :n_lag], sums[k, n_lag], sums[k, n_lag + 1])
        rh[k] = max(_rhat(sums[k, 0], sums[k, n_lag], sums[k, n_lag + 1], m, n),
                    _rhat(folded[k, 0], folded[k, 1], folded[k, 2], int(round(folded[k, 3])), n))
    return ess, rh


def _ess_from_sums_truncated(m, n, acov_sum, mean_sum, mean_sq_sum):
    """Geyer's sequences on the first n_lag lags only (the caller chose n_lag beyond the point where the
    paired autocorrelations turn negative)."""
    n_lag = len(acov_sum)
    acov_mean = np.asarray(acov_sum) / m
    mean_var = acov_mean[0] * n / (n - 1.0)
    var_plus = mean_var * (n - 1.0) / n
    if m > 1:
        var_plus += (mean_sq_sum - mean_sum ** 2 / m) / (m - 1.0)
    rho = 1.0 - (mean_var - acov_mean) / var_plus
    rho[0] = 1.0
    tau = -1.0
    prev = None
    t = 0
    while t + 1 < n_lag:
        pair = rho[t] + rho[t + 1]
        if pair < 0:
            break
        if prev is not None and pair > prev:
            pair = prev                                # initial monotone sequence
        tau += 2.0 * pair
        prev = pair
        t += 2
    tau = max(tau, 1.0 / np.log10(m * n))
    return m * n / tau
