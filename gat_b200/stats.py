"""Multiple-testing correction of a run's p-values.

Host numpy over at most tracks x annotations x counters numbers, executed once per run by
`outputResults` (reference call site: gat/IO.py:466-477).  Behaviour follows gat/Stats.py
(`adjustPValues` :192-258 = R's p.adjust; `computeQValues` :26-160 = Storey's q-value), written
vectorised; tests/test_host_logic.py pins both against golden vectors produced by the reference.
"""
import numpy as np


class FDRResult(object):
    """result record of computeQValues: qvalues, pvalues, pi0, vlambda, fdr_level, passed"""
    pass


def _step_up(p, n, scale):
    """min-accumulate of scale*p taken from the largest p downwards, mapped back to input order"""
    order = np.argsort(1 - p)
    back = np.argsort(order)
    return np.minimum.accumulate(scale * p[order])[back]


def adjustPValues(pvalues, method="fdr", n=None):
    """adjusted p-values: bonferroni, holm, hochberg, BH (alias fdr), BY or none."""
    p = np.array(pvalues, dtype=np.float64)
    count = len(p)
    n = count if n is None else n
    assert n <= count
    method = {"fdr": "BH"}.get(method, method)
    if n <= 1:
        return p
    if method == "hommel":
        if n != 2:
            raise NotImplementedError("hommel method not fully implemented")
        method = "hochberg"                       # identical for two tests
    descending_rank = np.arange(count, 0, -1)     # count .. 1 along the descending-p order
    if method == "none":
        adjusted = p
    elif method == "bonferroni":
        adjusted = n * p
    elif method == "holm":
        order = np.argsort(p)
        back = np.argsort(order)
        adjusted = np.maximum.accumulate((n - np.arange(count)) * p[order])[back]
    elif method == "hochberg":
        adjusted = _step_up(p, n, n - (descending_rank - 1))
    elif method == "BH":
        adjusted = _step_up(p, n, float(n) / descending_rank)
    elif method == "BY":
        harmonic = np.sum(1.0 / np.arange(1, n + 1))
        adjusted = _step_up(p, n, harmonic * float(n) / descending_rank)
    else:
        raise ValueError("unknown method '%s'" % method)
    return np.minimum(adjusted, 1.0)


def _estimate_pi0(p, vlambda, pi0_method, smooth_df, smooth_log_pi0):
    if isinstance(vlambda, float):
        vlambda = (vlambda,)
    nl = len(vlambda)
    if 1 < nl < 4:
        raise ValueError(" if length of vlambda greater than 1, you need at least 4 values.")
    if nl > 1 and (min(vlambda) < 0 or max(vlambda) >= 1):
        raise ValueError("vlambda must be within [0, 1).")
    if nl == 1:
        lam = vlambda[0]
        if lam < 0 or lam >= 1:
            raise ValueError("vlambda must be within [0, 1).")
        return min(np.mean(p >= lam) / (1.0 - lam), 1.0), lam
    lams = np.asarray(vlambda, dtype=np.float64)
    estimates = np.array([np.mean(p >= lam) / (1.0 - lam) for lam in lams])
    if pi0_method == "smoother":
        import scipy.interpolate
        y = np.log(estimates) if smooth_log_pi0 else estimates
        spline = scipy.interpolate.splrep(lams, y, k=smooth_df, s=10000)
        pi0 = scipy.interpolate.splev(max(lams), spline)
        if smooth_log_pi0:
            pi0 = np.exp(pi0)
    elif pi0_method == "bootstrap":
        floor = estimates.min()
        mse = np.zeros(nl)
        m = len(p)
        for _ in range(100):
            resampled = p[np.random.randint(0, m, m)]
            boot = np.array([np.mean(resampled > lam) / (1.0 - lam) for lam in lams])
            mse += (boot - floor) ** 2
        pi0 = estimates[mse == mse.min()].min()
    else:
        raise ValueError("'pi0_method' must be one of 'smoother' or 'bootstrap'.")
    return min(pi0, 1.0), vlambda


def computeQValues(pvalues, vlambda=None, pi0_method="smoother", fdr_level=None, robust=False,
                   smooth_df=3, smooth_log_pi0=False, pi0=None):
    """Storey-Tibshirani q-values with pi0 from a smoothing spline or bootstrap over lambda."""
    if min(pvalues) < 0 or max(pvalues) > 1:
        raise ValueError("p-values out of range")
    p = np.array(pvalues, dtype=np.float64)
    m = len(p)
    if vlambda is None:
        vlambda = np.arange(0, 0.95, 0.05)
    if pi0 is None:
        pi0, vlambda = _estimate_pi0(p, vlambda, pi0_method, smooth_df, smooth_log_pi0)
    if pi0 <= 0:
        raise ValueError("The estimated pi0 <= 0 (%f). Check that you have valid p-values "
                         "or use another vlambda method." % pi0)
    if fdr_level is not None and (fdr_level <= 0 or fdr_level > 1):
        raise ValueError("'fdr_level' must be within (0, 1].")

    ascending = np.argsort(p)
    sorted_p = p[ascending]
    n_le = np.searchsorted(sorted_p, p, side="right")        # observations <= p[i]
    q = p * pi0 * m / n_le
    if robust:
        q /= (1.0 - (1.0 - p) ** m)
    # cap at 1 and make monotone in p
    q_sorted = np.minimum.accumulate(np.minimum(q[ascending], 1.0)[::-1])[::-1]
    q = np.empty(m)
    q[ascending] = q_sorted

    result = FDRResult()
    result.qvalues = q
    result.passed = [bool(x <= fdr_level) for x in q] if fdr_level is not None else [False] * m
    result.pvalues = p
    result.pi0 = pi0
    result.vlambda = vlambda
    result.fdr_level = fdr_level
    return result
