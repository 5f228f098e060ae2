"""Exact assignment on the BASELINE config 1 instance (8gaussians -> 2moons, batch 256): sequential-step counts of the
shipped algorithm (row + column reduction, greedy start, shortest augmenting paths: what csrc/assign.cu runs) and of
three alternatives that were considered and rejected.  One "step" / "scan" / "round" is one
dependent block-wide operation on the GPU (~0.33 us).

    python scripts/sim/assignment_steps.py
"""
import os
import sys

import numpy as np
import torch
from scipy.optimize import linear_sum_assignment

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import eight_gaussians, two_moons  # noqa: E402


def c1(seed=10, n=256):
    gen = torch.Generator().manual_seed(seed)
    a, b = eight_gaussians(n, gen), two_moons(n, gen)
    M = (torch.cdist(a, b) ** 2).float().numpy().astype(np.float64)
    return M

def init_current(C):
    n = C.shape[0]
    u = C.min(1); pref = C.argmin(1)
    R = C - u[:, None]
    v = R.min(0); arow = R.argmin(0)
    r4c = -np.ones(n, int); c4r = -np.ones(n, int)
    for i in range(n):
        j = pref[i]
        if r4c[j] < 0: r4c[j] = i; c4r[i] = j
    for j in range(n):
        i = arow[j]
        if r4c[j] < 0 and c4r[i] < 0: r4c[j] = i; c4r[i] = j
    return u, v, r4c, c4r

def sap(C, u, v, r4c, c4r, order=None):
    n = C.shape[0]
    steps = 0; naug = 0
    rows = range(n) if order is None else order
    for cur in rows:
        if c4r[cur] >= 0: continue
        shortest = np.full(n, np.inf); path = -np.ones(n, int); sc = np.zeros(n, bool)
        i = cur; sink = -1; minval = 0.0; sr = []
        while sink < 0:
            sr.append(i)
            r = minval + C[i] - u[i] - v
            upd = (~sc) & (r < shortest)
            shortest[upd] = r[upd]; path[upd] = i
            cand = np.where(sc, np.inf, shortest)
            m = cand.min()
            js = np.flatnonzero(cand == m)
            fj = [j for j in js if r4c[j] < 0]
            j = fj[0] if fj else js[0]
            steps += 1
            minval = m; sc[j] = True
            if r4c[j] < 0: sink = j
            else: i = r4c[j]
        for ii in sr:
            if ii == cur: u[ii] += minval
            else: u[ii] += minval - shortest[c4r[ii]]
        v[sc] -= minval - shortest[sc]
        j = sink
        while True:
            ii = path[j]; r4c[j] = ii; jp = c4r[ii]; c4r[ii] = j; j = jp
            if ii == cur: break
        naug += 1
    return steps, naug

def arr(C, u, v, r4c, c4r, passes=2, cap=None):
    """LAPJV augmenting row reduction. v duals (column), returns number of row scans."""
    n = C.shape[0]
    scans = 0
    free = [i for i in range(n) if c4r[i] < 0]
    for _ in range(passes):
        k = 0; prev_free = free; free = []
        nfree0 = len(prev_free)
        budget = cap if cap else 10**9
        while k < len(prev_free):
            i = prev_free[k]; k += 1
            h = C[i] - v
            j1 = int(h.argmin()); u1 = h[j1]
            h2 = h.copy(); h2[j1] = np.inf
            j2 = int(h2.argmin()); u2 = h2[j2]
            scans += 1
            i0 = r4c[j1]
            if u1 < u2 and scans < budget:
                v[j1] -= (u2 - u1)
            elif i0 >= 0:
                j1 = j2; i0 = r4c[j1]
            r4c[j1] = i; c4r[i] = j1
            if i0 >= 0:
                c4r[i0] = -1
                if u1 < u2 and scans < budget:
                    k -= 1; prev_free[k] = i0      # process immediately
                else:
                    free.append(i0)
    # recompute u so that duals are feasible & matched tight
    R = C - v[None, :]
    u[:] = R.min(1)
    # matched edges must be tight: check
    for i in range(n):
        if c4r[i] >= 0 and abs(R[i, c4r[i]] - u[i]) > 1e-12 * max(1, abs(u[i])):
            # not tight -> unassign
            r4c[c4r[i]] = -1; c4r[i] = -1
    return scans

def auction(C, eps_list, keep=True, v0=None, maxpar=10**9):
    n = C.shape[0]
    p = np.zeros(n) if v0 is None else -v0.copy()
    log = []
    r4c = -np.ones(n, int); c4r = -np.ones(n, int)
    for eps in eps_list:
        if keep:
            # keep pairs satisfying eps-CS under the new eps
            V = C + p[None, :]
            best = V.min(1)
            for i in range(n):
                j = c4r[i]
                if j >= 0 and V[i, j] > best[i] + eps:
                    c4r[i] = -1; r4c[j] = -1
        else:
            r4c[:] = -1; c4r[:] = -1
        rounds = 0; bids = 0; wr = 0
        while (c4r < 0).any():
            free = np.flatnonzero(c4r < 0)[:maxpar]
            V = C[free] + p[None, :]
            j1 = V.argmin(1); best = V[np.arange(len(free)), j1]
            V[np.arange(len(free)), j1] = np.inf
            second = V.min(1)
            inc = second - best + eps
            rounds += 1; bids += len(free); wr += (len(free) + 31) // 32
            for j in np.unique(j1):
                who = np.flatnonzero(j1 == j)
                w = who[np.argmax(inc[who])]
                i = free[w]
                if r4c[j] >= 0: c4r[r4c[j]] = -1
                r4c[j] = i; c4r[i] = j
                p[j] += inc[w]
        log.append((rounds, bids, wr))
    return p, r4c, c4r, log

def warm_from_prices(C, p):
    n = C.shape[0]
    v = -p
    R = C - v[None, :]
    u = R.min(1); pref = R.argmin(1)
    r4c = -np.ones(n, int); c4r = -np.ones(n, int)
    for i in range(n):
        j = pref[i]
        if r4c[j] < 0: r4c[j] = i; c4r[i] = j
    return u, v.copy(), r4c, c4r

def bf_phase(C, u, v, r4c, c4r):
    """multi-source Bellman-Ford over alternating paths; returns rounds, number of augmentations."""
    n = C.shape[0]
    free_rows = np.flatnonzero(c4r < 0)
    R = C - u[:, None] - v[None, :]          # reduced costs >= 0
    # round 0: from free rows
    k = R[free_rows].argmin(0)
    d = R[free_rows, np.arange(n)][k, np.arange(n)] if False else R[free_rows].min(0)
    pred = free_rows[k]
    rounds = 1
    matched_rows = np.flatnonzero(c4r >= 0)
    while True:
        if len(matched_rows) == 0: break
        drow = d[c4r[matched_rows]]           # label of matched rows = label of their column
        T = drow[:, None] + R[matched_rows]
        k = T.argmin(0); dn = T[k, np.arange(n)]
        upd = dn < d
        rounds += 1
        if not upd.any(): break
        d = np.where(upd, dn, d); pred = np.where(upd, matched_rows[k], pred)
    # potentials: rows: drow_i = d[c4r[i]] (matched), 0 (free)
    drow = np.zeros(n); drow[matched_rows] = d[c4r[matched_rows]]
    u_new = u - drow; v_new = v + d            # c - u_new - v_new = R + drow_i - d_j >= 0
    # disjoint augmenting paths via pred tree, free columns in order of d
    free_cols = np.flatnonzero(r4c < 0)
    free_cols = free_cols[np.argsort(d[free_cols], kind='stable')]
    used_row = np.zeros(n, bool); used_col = np.zeros(n, bool)
    naug = 0
    for jc in free_cols:
        # walk
        pathc = []; pathr = []; j = jc; ok = True
        while True:
            if used_col[j]: ok = False; break
            i = pred[j]
            if used_row[i]: ok = False; break
            pathc.append(j); pathr.append(i)
            if c4r[i] < 0: break
            j = c4r[i]
        if not ok: continue
        for j, i in zip(pathc, pathr):
            used_col[j] = True; used_row[i] = True
        # augment
        for j, i in zip(pathc, pathr):
            r4c[j] = i
        for j, i in zip(pathc, pathr):
            c4r[i] = j
        naug += 1
    # fix r4c for columns that lost their row: handled since path columns each get new row; 
    return u_new, v_new, rounds, naug


if __name__ == '__main__':
    for seed in (10, 11, 12):
        C = c1(seed)
        ri, ci = linear_sum_assignment(C)
        u, v, r4c, c4r = init_current(C)
        nfree = int((c4r < 0).sum())
        steps, naug = sap(C, u.copy(), v.copy(), r4c.copy(), c4r.copy())
        print(f"seed {seed}: shipped algorithm: {nfree} free rows after the greedy start, {steps} Dijkstra steps")
        u2, v2, r, c = u.copy(), v.copy(), r4c.copy(), c4r.copy()
        scans = arr(C, u2, v2, r, c, 1)
        s1, _ = sap(C, u2, v2, r, c)
        print(f"   + JV augmenting row reduction (1 pass): {scans} row scans, then {s1} Dijkstra steps (sigma ok: {np.array_equal(c, ci)})")
        cmax = C.max()
        p, r, c, log = auction(C, [cmax / 4 ** k for k in range(1, 12)], keep=False)
        u3, v3, r3, c3 = warm_from_prices(C, p)
        s2, _ = sap(C, u3, v3, r3, c3)
        print(f"   + epsilon-scaling Jacobi auction warm start: {sum(l[0] for l in log)} rounds, then {s2} Dijkstra steps (sigma ok: {np.array_equal(c3, ci)})")
        u4, v4, r4, c4 = init_current(C)
        rounds = phases = 0
        while (c4 < 0).any():
            u4, v4, rr, na = bf_phase(C, u4, v4, r4, c4)
            rounds += rr; phases += 1
        print(f"   multi-source Bellman-Ford phases, all disjoint shortest paths augmented: {phases} phases, {rounds} rounds (sigma ok: {np.array_equal(c4, ci)})")
