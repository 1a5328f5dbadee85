"""Rooted trees and Runge–Kutta order conditions (Butcher): used by tools/derive_vern6.py and
tests/test_vern6_order_conditions.py.  Pure Python; works with floats, Fractions or mpmath numbers."""
from functools import lru_cache


@lru_cache(maxsize=None)
def trees(order):
    """All rooted trees with `order` vertices, each as a sorted tuple of its root's subtrees."""
    if order == 1:
        return ((),)
    out = set()

    def parts(n, maxpart):  # partitions of n into parts <= maxpart
        if n == 0:
            yield ()
            return
        for p in range(min(n, maxpart), 0, -1):
            for rest in parts(n - p, p):
                yield (p,) + rest

    def combos(sizes, idx=0):
        if idx == len(sizes):
            yield ()
            return
        for t in trees(sizes[idx]):
            for rest in combos(sizes, idx + 1):
                yield (t,) + rest

    for part in parts(order - 1, order - 1):
        for kids in combos(part):
            out.add(tuple(sorted(kids)))
    return tuple(sorted(out))


def rho(t):
    return 1 + sum(rho(k) for k in t)


def gamma(t):
    g = rho(t)
    for k in t:
        g *= gamma(k)
    return g


def trees_upto(order):
    return [t for q in range(1, order + 1) for t in trees(q)]


def stage_weights(A, t, one=1):
    """Φ_i(t) for every stage i of the (strictly lower triangular) matrix A (list of rows)."""
    s = len(A)
    if not t:
        return [one] * s
    out = [one] * s
    for kid in t:
        phi = stage_weights(A, kid, one)
        for i in range(s):
            acc = 0 * one
            for j in range(len(A[i])):
                if A[i][j] != 0:
                    acc += A[i][j] * phi[j]
            out[i] = out[i] * acc
    return out


def residuals(A, b, order, one=1, theta=None):
    """Σ b_i Φ_i(t) − θ^ρ(t)/γ(t) for all trees up to `order` (θ = 1: the discrete method)."""
    res = []
    for t in trees_upto(order):
        phi = stage_weights(A, t, one)
        lhs = sum(bi * p for bi, p in zip(b, phi) if bi != 0)
        rhs = one / gamma(t) if theta is None else theta ** rho(t) / gamma(t)
        res.append(lhs - rhs)
    return res
