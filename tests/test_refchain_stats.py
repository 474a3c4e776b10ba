"""CPU: the statistics the GPU posterior tests rest on (tests/refchain.py), tried on synthetic chains whose truth is known:
batch-means standard errors of autocorrelated series, and the two-group criterion of the configs[3] tests — 3 standard
errors of the difference of the group means (larger of the between-chain and the batch-means estimate) + 2 %."""
import numpy as np
from scipy.signal import lfilter

import refchain as rc


def ar1(rng, n, rho, mean, sd):
    """stationary AR(1) series with the given marginal mean and standard deviation"""
    e = rng.normal(0.0, sd * np.sqrt(1.0 - rho * rho), n)
    e[0] = rng.normal(0.0, sd)
    return mean + lfilter([1.0], [1.0, -rho], e)      # x[i] = rho * x[i-1] + e[i]


def test_batch_means_error_covers_the_truth_for_autocorrelated_chains():
    rng = np.random.default_rng(5)
    rho, sd, n = 0.95, 1.0, 24000
    true_se = sd * np.sqrt((1 + rho) / (1 - rho) / n)
    z, est = [], []
    for _ in range(60):
        x = ar1(rng, n, rho, 10.0, sd)
        se = rc.batch_se(x)
        est.append(se)
        z.append((x.mean() - 10.0) / se)
    assert 0.75 * true_se < np.mean(est) < 1.25 * true_se
    assert np.mean(np.abs(z) < 3.0) > 0.93          # 20 batches: Student t with 19 degrees of freedom


def accept(diff, se, ref):
    return bool(np.all(np.abs(diff) < 3.0 * se + 0.02 * np.abs(ref)))


def test_two_group_criterion_accepts_equal_posteriors_and_rejects_a_shifted_one():
    """three reference chains against two or three device chains, 12 parameters of which some mix slowly"""
    rng = np.random.default_rng(11)
    K, n = 12, 12000
    rhos = np.r_[np.full(8, 0.9), np.full(4, 0.999)]      # the last four: integrated autocorrelation time ~ 2000 iterations
    means = np.linspace(1.0, 5.0, K)

    def chain(shift=0.0):
        return np.stack([ar1(rng, n, rhos[k], means[k] * (1.0 + (shift if k == 3 else 0.0)), 0.05 * means[k]) for k in range(K)], 1)

    ok_same = ok_shift = 0
    trials = 40
    for _ in range(trials):
        a = [chain() for _ in range(3)]
        b = [chain() for _ in range(2)]
        diff, se, ma, _ = rc.group_difference(a, b)
        ok_same += accept(diff, se, ma.mean(0))
        c = [chain(shift=0.06) for _ in range(2)]        # one well-mixing parameter 6 % off
        diff, se, ma, _ = rc.group_difference(a, c)
        ok_shift += accept(diff, se, ma.mean(0))
    assert ok_same >= trials - 3, ok_same                  # equal posteriors pass (slowly mixing parameters included)
    assert ok_shift <= 2, ok_shift                         # a 6 % bias in one parameter does not


def test_between_chain_error_alone_is_too_noisy_for_five_chains():
    """why group_difference takes the larger of two estimates: with 3 + 2 chains the pooled between-chain error has 3
    degrees of freedom, and by chance it comes out several times too small"""
    rng = np.random.default_rng(2)
    n, trials, small = 8000, 300, 0
    for _ in range(trials):
        a = [ar1(rng, n, 0.9, 1.0, 0.05)[:, None] for _ in range(3)]
        b = [ar1(rng, n, 0.9, 1.0, 0.05)[:, None] for _ in range(2)]
        between = rc.pooled_between_chain_se(np.array([c.mean(0) for c in a]), np.array([c.mean(0) for c in b]))[0]
        _, se, _, _ = rc.group_difference(a, b)
        small += between < 0.4 * se[0]
    assert small > trials // 20
