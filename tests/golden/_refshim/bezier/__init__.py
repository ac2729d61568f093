"""Stand-in for the un-vendored `bezier` PyPI package (Apache-2; the reference
pins no version) -- TEST INFRASTRUCTURE ONLY, used by tests/golden/make_golden.py.

Only `bezier.Curve(nodes, degree).evaluate_multi(s_vals)` is needed
(reference dynamics_and_models.py:616-618, 649-651, 684-686).  The published
algorithm restated here is bezier's `evaluate_multi_barycentric` /
`evaluate_multi_vectorized` (bezier/hazmat/curve_helpers.py, releases 2020.1 ..
2024.6): a Horner-like scheme in barycentric weights (1-s, s), all in float64:

    result  = (1-s) * P0
    for k = 1 .. degree-1:
        s_pow *= s ; binom = binom * (degree-k+1) / k
        result += binom * s_pow * Pk
        result *= (1-s)
    result += s * s_pow * P_degree

Nodes are promoted losslessly to float64 first (Curve.__init__ ->
`lossless_to_float`).
"""
import numpy as np


class Curve(object):
    def __init__(self, nodes, degree, copy=True, verify=True):
        self._nodes = np.asfortranarray(np.asarray(nodes).astype(np.float64))
        self._degree = int(degree)
        assert self._nodes.shape[1] == self._degree + 1

    @property
    def nodes(self):
        return self._nodes.copy()

    def evaluate_multi(self, s_vals):
        s = np.asarray(s_vals, dtype=np.float64)
        lambda1 = (1.0 - s)[np.newaxis, :]
        lambda2 = s[np.newaxis, :]
        nodes = self._nodes
        degree = self._degree
        result = np.zeros((nodes.shape[0], s.shape[0]), order='F')
        result += lambda1 * nodes[:, [0]]
        binom_val = 1.0
        lambda2_pow = np.ones((1, s.shape[0]), order='F')
        for index in range(1, degree):
            lambda2_pow *= lambda2
            binom_val = (binom_val * (degree - index + 1)) / index
            result += binom_val * lambda2_pow * nodes[:, [index]]
            result *= lambda1
        result += lambda2 * lambda2_pow * nodes[:, [degree]]
        return result
