"""numpy model of the odd-even one-sided Jacobi of qsp_jacobi.cuh (same rotation, same stop rule) used to choose the
row placement before every sweep ("Norm ordering").  Prints the sweeps to convergence of 128 x 128
complex matrices (full rank, and graded over four decades) for the placements that were tried:

    none       rows stay where the last sweep left them (round 1)
    desc       decreasing norm along the line
    asc        increasing norm along the line
    fold       largest next to smallest
    ring       decreasing along the ring of the schedule (0, 2, ..., n-2, n-1, ..., 3, 1)
    ring_rev   decreasing along the reversed ring (1, 3, ..., n-1, n-2, ..., 2, 0)   <- the kernel's choice

    python scripts/sim_jacobi_ordering.py [matrices per kind]        (a few minutes)
"""
import sys

import numpy as np

QUAD = 3e-8


def gen(rng, kind, n=128):
    if kind == "gauss":
        return rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    u, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    v, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    return (u * 10.0 ** (-4 * np.sort(rng.random(n)))) @ v.conj().T


def place(order, n, strat):
    """positions (index into the line) of the rows listed in `order` (decreasing norm)"""
    ring = np.concatenate([np.arange(0, n, 2), np.arange(n - 1 if n % 2 == 0 else n - 2, 0, -2)])
    if strat == "desc":
        return np.arange(n)
    if strat == "asc":
        return np.arange(n)[::-1]
    if strat == "fold":
        pos = np.empty(n, int)
        pos[: n // 2] = np.arange(0, n, 2)[: n // 2]
        pos[n // 2:] = np.arange(1, n, 2)[::-1][: n - n // 2]
        return pos
    if strat == "ring":
        return ring
    if strat == "ring_rev":
        return ring[::-1]
    raise ValueError(strat)


def sweeps(W, strat, tolf=4e-16, maxs=60):
    n, m = W.shape
    W = W.copy()
    tol2 = (tolf * np.sqrt(m)) ** 2
    for sw in range(1, maxs + 1):
        if strat != "none":
            order = np.argsort(-np.linalg.norm(W, axis=1), kind="stable")
            Wn = np.empty_like(W)
            Wn[place(order, n, strat)] = W[order]
            W = Wn
        flag = False
        for r in range(n):
            i0 = np.arange(r & 1, n - 1, 2)
            i1 = i0 + 1
            x, y = W[i0], W[i1]
            al, be = (np.abs(x) ** 2).sum(1), (np.abs(y) ** 2).sum(1)
            g = (x * np.conj(y)).sum(1)
            g2, ab = np.abs(g) ** 2, al * be
            rot = (ab > 1e-280) & (g2 > tol2 * ab)
            flag |= bool(np.any(rot & (g2 > QUAD ** 2 * ab)))
            d = be - al
            h = np.sqrt(d * d + 4 * g2)
            u = h + np.abs(d)
            with np.errstate(all="ignore"):
                sw_ = 1 / np.sqrt(0.5 * h * u)
            cs = np.where(rot, 0.5 * u * sw_, 1.0)
            se = np.where(rot, np.where(d >= 0, sw_, -sw_), 0.0) * g
            W[i0], W[i1] = np.conj(se)[:, None] * x + cs[:, None] * y, cs[:, None] * x - se[:, None] * y  # trade places
        if not flag:
            return sw
    return maxs


if __name__ == "__main__":
    nm = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    rng = np.random.default_rng(5)
    for kind in ("gauss", "graded4"):
        mats = [gen(rng, kind) for _ in range(nm)]
        for strat in ("none", "desc", "asc", "fold", "ring", "ring_rev"):
            v = [sweeps(W, strat) for W in mats]
            print(f"{kind:8s} {strat:9s} mean {np.mean(v):5.2f}  {v}", flush=True)
