"""Host-side significance helpers behind ``Connectivity.delay`` / ``group_delay``.

These are O(n_frequencies) bookkeeping on top of the device-computed coherency (p-values, a multiple-comparison
threshold, connected runs along the frequency axis): index work, not a hot path, so they stay in NumPy/SciPy exactly
like the reference (statistics.py:21-59, 147-203, 206-247, 250-288; connectivity.py:2102-2243).

Faithfulness note: the reference calls ``coherence_fisher_z_transform(coherency, n_obs)`` with its defaults
``coherency2=0, n_obs2=0`` (connectivity.py:2227).  The bias of the second group is then 1 / (2*0 - 2) = -0.5 and the
normalisation sqrt(bias1 + bias2) is the square root of a negative number: every z-score is NaN, no frequency is ever
significant, and ``delay`` / ``group_delay`` return their all-masked results (verified by running the live reference,
tests/golden/make_golden.py: delay_section).  The same arithmetic is kept here so that the outputs are identical.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import label
from scipy.special import ndtr


def coherence_bias(n_observations):
    """statistics.py:250-288: 1 / (degrees of freedom - 2) with 2 degrees of freedom per observation."""
    return 1.0 / (2 * n_observations - 2)


def fisher_z(coherency, n_obs, coherency2=0, n_obs2=0):
    """statistics.py:147-203: bias-corrected arctanh of the coherence magnitude(s), difference over pooled bias."""
    with np.errstate(invalid="ignore", divide="ignore"):
        mag1 = np.abs(coherency)
        mag1[mag1 >= 1] = 1 - np.finfo(float).eps
        mag2 = np.array(np.abs(coherency2))
        mag2[mag2 >= 1] = 1 - np.finfo(float).eps
        b1, b2 = coherence_bias(n_obs), coherence_bias(n_obs2)
        return ((np.arctanh(mag1) - b1) - (np.arctanh(mag2) - b2)) / np.sqrt(b1 + b2)


def upper_tail_p(z):
    """statistics.py:206-247: 1 - Phi(z)."""
    return 1 - ndtr(z)


def benjamini_hochberg(p_values, alpha=0.05):
    """statistics.py:21-59: one family over the flattened array; reject p <= largest sorted p under the BH line."""
    p_values = np.array(p_values)
    line = np.linspace(0, alpha, num=p_values.size + 1, endpoint=True)[1:]
    ordered = np.sort(p_values.flatten())
    below = np.where(ordered <= line)[0]
    threshold = ordered[int(below.max())] if below.size else -1
    return p_values <= threshold


def bonferroni(p_values, alpha=0.05):
    """statistics.py:62-98."""
    p_values = np.array(p_values)
    return p_values <= alpha / p_values.size


def _largest_run(flags):
    """connectivity.py:2102-2129: keep only the longest run of consecutive True values (first one on ties)."""
    runs, _ = label(flags)
    ids, counts = np.unique(runs, return_counts=True)
    if np.all(ids == 0):
        return np.zeros(flags.shape, dtype=bool)
    counts[0] = 0
    return runs == ids[np.argmax(counts)]


def _independent_run(flags, frequency_step, min_group_size):
    """connectivity.py:2132-2182: every ``frequency_step``-th point of the largest run, or nothing if too few."""
    flags = _largest_run(flags)
    hits = flags.nonzero()[0]
    flags = np.isin(np.arange(len(flags)), hits[0:len(hits):frequency_step])
    if flags.sum() < min_group_size:
        flags[:] = False
    return flags


def significant_frequencies(coherency, n_obs, frequency_step=1, significance_threshold=0.05, min_group_size=3,
                            multiple_comparisons_method="Benjamini_Hochberg_procedure"):
    """connectivity.py:2185-2243 for an array (..., n_frequencies, n_pairs): bool mask of the same shape."""
    adjust = {"Benjamini_Hochberg_procedure": benjamini_hochberg, "Bonferroni_correction": bonferroni}
    keep = adjust[multiple_comparisons_method](upper_tail_p(fisher_z(coherency, n_obs)), alpha=significance_threshold)
    return np.apply_along_axis(_independent_run, -2, keep, frequency_step, min_group_size)
