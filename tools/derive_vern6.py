#!/usr/bin/env python
"""Derives Verner's "most efficient" 9-stage 6(5) pair (OrdinaryDiffEq's `Vern6`, the integrator of the reference's
`propagate`, RayTracer.jl:383) and a 6th-order continuous extension on three extra stages, and writes them as C
headers for the device code and for the oracle:

    adiabatic_raytracer_b200/csrc/art_vern6.cuh      oracle/vern6.hpp

The tableau is not on disk (OrdinaryDiffEq is an un-vendored dependency of the reference) and there is no network.
What is known:
  * the nodes  c = 0, 3/50, 1439/15000, 1439/10000, 4973/10000, 389/400, 1999/2000, 1, 1   (exact rationals),
  * the sparsity of Verner's design:  a_i2 = 0 for i >= 4,  b_2 = b_3 = b_9 = 0,  FSAL (a_9j = b_j),
    embedded weights with  bh_2 = bh_3 = bh_7 = 0,
  * the coefficients as 16-digit decimals recalled from OrdinaryDiffEq's verner_tableaus.jl (START below).
The recalled decimals are NOT trusted: they are only the starting point of a Newton iteration, in 70-digit
arithmetic, on the full set of conditions — 7 row sums, all 37 order conditions of order <= 6 for b and all 17 of
order <= 5 for bh.  That system has full column rank (smallest singular value 8e-8), i.e. the pair is UNIQUELY
determined by nodes + sparsity + order conditions; the iteration converges to residual 1e-66 and moves the start
by 8e-12 in one coefficient (a_54, a slip in the 11th digit of the recollection) and < 1e-16 in every other one.
bh_1, bh_5, bh_6, bh_8 agree with Verner's published 34-digit values to every digit.

Dense output.  OrdinaryDiffEq's Vern6 uses a lazy 6th-order interpolant on three extra stages (k10..k12), which the
reference's ContinuousCallback evaluates at 50 points of every accepted step.  Only the first extra stage is
recalled (c10 = 1/2 and its row; it satisfies all order <= 5 conditions of an internal stage at c10 to 1e-17, i.e.
it is the 5th-order dense output at mid-step); rows 11, 12 and the weight polynomials are NOT recalled.  They are
constructed here — this project's own specification, identical in oracle and device code:
  * c11 = 7/25, c12 = 19/100; rows 11 and 12 are the minimum-norm solutions of the order <= 5 internal-stage
    conditions  sum_j a_ij Phi_j(t) = c_i^rho(t) / gamma(t)   over the stages 1, 4..i-1;
  * the weights  b_i(theta) = sum_{m=1..6} R[i][m] theta^m  solve  sum_i b_i(theta) Phi_i(t) = theta^rho(t)/gamma(t)
    for all 37 trees of order <= 6, power by power (rank 10 = number of unknown weights: unique).
Its principal error norm over theta in (0,1) is 2e-4 against 2.6e-6 of the discrete formula at theta = 1; this
floor is set by stage 10 (any choice of c11, c12 and of the free row parameters gives 1.5e-4 .. 5e-4), so
Verner's own interpolant cannot be much different in accuracy; the dense-output error at the reference's
tolerances is measured in tests/test_vern6_order_conditions.py.

Usage:  python tools/derive_vern6.py            (rewrites the two headers; ~1 minute)
"""
import os
import sys

import mpmath as mp
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from rk_trees import gamma, residuals, rho, stage_weights, trees_upto  # noqa: E402

mp.mp.dps = 70
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def Q(p, q):
    return mp.mpf(p) / mp.mpf(q)


C = [mp.mpf(0), Q(3, 50), Q(1439, 15000), Q(1439, 10000), Q(4973, 10000), Q(389, 400), Q(1999, 2000), mp.mpf(1),
     mp.mpf(1), Q(1, 2), Q(7, 25), Q(19, 100)]

# recollection of OrdinaryDiffEq's Vern6Tableau (1-based stage indices) — starting point only
START = {(2, 1): 0.06, (3, 1): 0.019239962962962962, (3, 2): 0.07669337037037037, (4, 1): 0.035975, (4, 3): 0.107925,
         (5, 1): 1.3186834152331484, (5, 3): -5.042058063628562, (5, 4): 4.220674648403414,
         (6, 1): -41.872591664327516, (6, 3): 159.4325621631375, (6, 4): -122.11921356501003, (6, 5): 5.531743066200054,
         (7, 1): -54.430156935316504, (7, 3): 207.06725136501848, (7, 4): -158.61081378459, (7, 5): 6.991816585950242,
         (7, 6): -0.018597231062203234,
         (8, 1): -54.66374178728198, (8, 3): 207.95280625538936, (8, 4): -159.2889574744995, (8, 5): 7.018743740796944,
         (8, 6): -0.018338785905045722, (8, 7): -0.0005119484997882099,
         (9, 1): 0.03438957868357036, (9, 4): 0.2582624555633503, (9, 5): 0.4209371189673537, (9, 6): 4.40539646966931,
         (9, 7): -176.48311902429865, (9, 8): 172.36413340141507}
START_BH = {1: 0.0490996764838249, 4: 0.2251112229516524, 5: 0.4694682253029562, 6: 0.8065792249988868,
            8: -0.607119489177796, 9: 0.0568611394404757}
START_ROW10 = {1: 0.016524159013572806, 4: 0.3053128187514179, 5: 0.2071200938201979, 6: -1.293879140655123,
               7: 57.11988411588149, 8: -55.87979207510932, 9: 0.024830028297766014}


def lsq(J, rhs):
    return mp.lu_solve(J.T * J, J.T * rhs)


def solve_pair(start=None, start_bh=None, verbose=True):
    """Newton on nodes + sparsity + order conditions from the given starting coefficients ({(i, j): a_ij} with
    row 9 = b, {j: bh_j}; 1-based).  Returns the 9 x 9 matrix A and bh as mpmath numbers."""
    start = START if start is None else start
    start_bh = START_BH if start_bh is None else start_bh
    keys = sorted(START)
    bkeys = sorted(START_BH)
    x = [mp.mpf(start[k]) for k in keys] + [mp.mpf(start_bh[k]) for k in bkeys]
    say = print if verbose else (lambda *a, **k: None)

    def unpack(x):
        A = [[mp.mpf(0)] * 9 for _ in range(9)]
        for k, v in zip(keys, x):
            A[k[0] - 1][k[1] - 1] = v
        bh = [mp.mpf(0)] * 9
        for k, v in zip(bkeys, x[len(keys):]):
            bh[k - 1] = v
        return A, bh

    def F(x):
        A, bh = unpack(x)
        res = [sum(A[i]) - C[i] for i in range(1, 8)]
        res += residuals(A, A[8], 6, mp.mpf(1))
        res += residuals(A, bh, 5, mp.mpf(1))
        return res

    eps = mp.mpf(10) ** -35
    for it in range(6):
        f0 = F(x)
        J = mp.matrix(len(f0), len(x))
        for j in range(len(x)):
            xp = list(x)
            xp[j] += eps
            fp = F(xp)
            for i in range(len(f0)):
                J[i, j] = (fp[i] - f0[i]) / eps
        if it == 0:
            sv = np.linalg.svd(np.array(J.tolist(), dtype=float), compute_uv=False)
            say(f"pair: {len(f0)} conditions, {len(x)} unknowns, smallest singular value {sv[-1]:.2e} (full rank)")
            assert sv[-1] > 1e-9, "the pair must be uniquely determined" 
        dx = lsq(J, mp.matrix([-v for v in f0]))
        x = [a + d for a, d in zip(x, dx)]
        say(f"  newton {it}: residual {mp.nstr(max(abs(v) for v in f0), 3)}, step {mp.nstr(max(abs(d) for d in dx), 3)}")
    A, bh = unpack(x)
    moved = {k: float(v) - start[k] for k, v in zip(keys, x) if float(v) != start[k]}
    say("  coefficients whose double differs from the start:", moved)
    assert max(abs(v) for v in F(x)) < mp.mpf(10) ** -50
    return A, bh


def internal_row(A, ci, free, start=None):
    """Row of an extra stage at node ci over the stages `free` (0-based): order <= 5 internal conditions; the
    solution nearest to `start` (minimum norm if None) when the conditions leave freedom."""
    T5 = trees_upto(5)
    n = len(A)
    M = mp.matrix(len(T5), len(free))
    rhs = mp.matrix(len(T5), 1)
    for r, t in enumerate(T5):
        phi = stage_weights(A, t, mp.mpf(1))
        for cidx, j in enumerate(free):
            M[r, cidx] = phi[j]
        rhs[r] = ci ** rho(t) / gamma(t)
    Mn = np.array(M.tolist(), dtype=float)
    U, S, Vt = np.linalg.svd(Mn)
    rank = int((S > 1e-11 * S[0]).sum())
    N = Vt[rank:]  # null space (float accuracy is enough: it only selects WHICH exact solution)
    x0 = mp.matrix([mp.mpf(0)] * len(free)) if start is None else mp.matrix([mp.mpf(v) for v in start])
    J = mp.matrix(len(T5) + len(N), len(free))
    b = mp.matrix(len(T5) + len(N), 1)
    for i in range(len(T5)):
        for j in range(len(free)):
            J[i, j] = M[i, j]
        b[i] = rhs[i] - sum(M[i, j] * x0[j] for j in range(len(free)))
    for k in range(len(N)):
        for j in range(len(free)):
            J[len(T5) + k, j] = mp.mpf(float(N[k, j]))
        b[len(T5) + k] = 0
    dx = lsq(J, b)
    x = [x0[j] + dx[j] for j in range(len(free))]
    res = max(abs(sum(M[i, j] * x[j] for j in range(len(free))) - rhs[i]) for i in range(len(T5)))
    print(f"  extra stage at c = {mp.nstr(ci, 6)}: rank {rank} of {len(free)} unknowns, residual {mp.nstr(res, 3)}, "
          f"moved from start by {mp.nstr(max(abs(d) for d in dx), 3)}")
    row = [mp.mpf(0)] * n
    for j, v in zip(free, x):
        row[j] = v
    return row


def weight_polynomials(A):
    T6 = trees_upto(6)
    idx = [0, 3, 4, 5, 6, 7, 8, 9, 10, 11]
    M = mp.matrix(len(T6), len(idx))
    for r, t in enumerate(T6):
        phi = stage_weights(A, t, mp.mpf(1))
        for cidx, j in enumerate(idx):
            M[r, cidx] = phi[j]
    sv = np.linalg.svd(np.array(M.tolist(), dtype=float), compute_uv=False)
    print(f"  weight polynomials: {len(T6)} conditions per power, {len(idx)} unknowns, singular values {sv[0]:.2e} .. {sv[-1]:.2e}")
    R = [[mp.mpf(0)] * 7 for _ in range(12)]
    worst = mp.mpf(0)
    for m in range(1, 7):
        rhs = mp.matrix([(mp.mpf(1) / gamma(t) if rho(t) == m else mp.mpf(0)) for t in T6])
        x = lsq(M, rhs)
        worst = max(worst, max(abs(v) for v in (M * x - rhs)))
        for j, v in zip(idx, x):
            R[j][m] = v
    print(f"  continuous order-6 conditions: residual {mp.nstr(worst, 3)}")
    return R


def fmt(v):
    return repr(float(v))


def emit(A, bh, R, path, device):
    lines = []
    w = lines.append
    w("// GENERATED by tools/derive_vern6.py — do not edit.  Verner's 9-stage 6(5) \"most efficient\" pair (OrdinaryDiffEq")
    w("// Vern6, RayTracer.jl:383) solved from its nodes, sparsity and order conditions in 70-digit arithmetic, and a")
    w("// 6th-order continuous extension on three extra stages (see the script's header for what is Verner's and what is")
    w("// this project's own construction).  Checked by tests/test_vern6_order_conditions.py.")
    w("#pragma once")
    if device:
        w("#if defined(__CUDACC__)")
        w("#define ART_V6_TABLE __device__ __constant__")
        w("#else")
        w("#define ART_V6_TABLE static const")
        w("#endif")
        w("namespace art {")
        q = "ART_V6_TABLE double"
    else:
        w("namespace orc {")
        w("namespace v6 {")
        q = "static const double"
    w("// nodes c_1..c_12 (stages 10..12: dense output only)")
    w(f"{q} V6_C[12] = {{" + ", ".join(fmt(c) for c in C) + "};")
    w("// a_ij, row-major 12 x 12 (row 9 = b: FSAL)")
    w(f"{q} V6_A[12][12] = {{")
    for i in range(12):
        w("    {" + ", ".join(fmt(A[i][j]) if j < len(A[i]) else "0.0" for j in range(12)) + "},")
    w("};")
    w("// error weights btilde_j = b_j - bh_j (5th-order embedded solution), stages 1..9")
    bt = [A[8][j] - bh[j] for j in range(9)]
    w(f"{q} V6_BT[9] = {{" + ", ".join(fmt(v) for v in bt) + "};")
    w("// dense output: u(theta) = u0 + h * sum_i (sum_{m=1..6} V6_R[i][m-1] theta^m) k_i")
    w(f"{q} V6_R[12][6] = {{")
    for i in range(12):
        w("    {" + ", ".join(fmt(R[i][m]) for m in range(1, 7)) + "},")
    w("};")
    # compile-time sparsity (the rolled stage loop skips exact zeros)
    w("// sparsity of V6_A as bit masks (bit j set: a_ij != 0); a constexpr function so that device code can fold it")
    masks = [sum(1 << j for j in range(12) if j < len(A[i]) and A[i][j] != 0) for i in range(12)]
    w("#if defined(__CUDACC__)")
    w("__host__ __device__")
    w("#endif")
    w("constexpr unsigned v6_mask(int i) {")
    w("  return " + " : ".join(f"i == {i} ? {hex(m)}u" for i, m in enumerate(masks)) + " : 0u;")
    w("}")
    if device:
        w("}  // namespace art")
    else:
        w("}  // namespace v6")
        w("}  // namespace orc")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", os.path.relpath(path, ROOT))


def main():
    A9, bh = solve_pair()
    A = [row + [mp.mpf(0)] * 3 for row in A9] + [[mp.mpf(0)] * 12 for _ in range(3)]
    print("dense output:")
    free10 = [0, 3, 4, 5, 6, 7, 8]
    A[9] = internal_row([r[:9] for r in A[:9]], C[9], free10, [START_ROW10[j + 1] for j in free10]) + [mp.mpf(0)] * 3
    A[10] = internal_row([r[:10] for r in A[:10]], C[10], [0, 3, 4, 5, 6, 7, 8, 9]) + [mp.mpf(0)] * 2
    A[11] = internal_row([r[:11] for r in A[:11]], C[11], [0, 3, 4, 5, 6, 7, 8, 9, 10]) + [mp.mpf(0)]
    for i in range(9, 12):
        assert abs(sum(A[i]) - C[i]) < mp.mpf(10) ** -40
    R = weight_polynomials(A)
    # sanity: theta = 1 reproduces b (and weights of the extra stages vanish there)
    for i in range(12):
        b1 = sum(R[i][m] for m in range(1, 7))
        target = A[8][i] if i < 9 else mp.mpf(0)
        assert abs(b1 - target) < mp.mpf(10) ** -40, (i, b1, target)
    emit(A, bh, R, os.path.join(ROOT, "adiabatic_raytracer_b200", "csrc", "art_vern6.cuh"), True)
    emit(A, bh, R, os.path.join(ROOT, "oracle", "vern6.hpp"), False)


if __name__ == "__main__":
    main()
