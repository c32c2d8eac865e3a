"""Synthetic second-quantised molecular Hamiltonians and sample sets (workload generation).

The reference obtains its Pauli dictionary from third-party packages that are not vendored and
not installed here (openfermion `jordan_wigner` / `get_molecular_hamiltonian`, pyscf integrals;
reference call sites nqs/nqs/applications/quantum_chemistry/molecular_data.py:11-13,56-66 and
run_pyscf.py:159-192).  Only the *output format* matters to the hot path
(`QubitOperator.terms`, pauli_observable.py:150-183), so this module produces operators of the
same shape from random integrals (SURVEY.md §8(d) "synthetic Hamiltonians"):

  H = c + sum_pq h1[p,q] a+_p a_q + 1/2 sum_pqrs G[p,q,r,s] a+_p a+_q a_r a_s   (spin-orbitals)

with spin-orbital index 2p+sigma, G[p,q,r,s] = (ps|qr), spin pattern (s,t,t,s), followed by a
vectorised Jordan-Wigner transform.  Qubit q lives at bit n-1-q of the packed index
(pauli_observable.py:162), i.e. qubit 0 is the most significant bit.

Internally every Pauli string is carried as O(x,z) = X^x Z^z (no Y phase): the coefficient of
O(x,z) is exactly the reference's table weight "coefficient * i^{#Y}" (pauli_observable.py:176-177),
and <x|O(xm,zm)|x'> = (-1)^{popcount(zm & x')} delta(x, x' xor xm).
"""
from __future__ import annotations

import numpy as np

U64 = np.uint64


def synthetic_integrals(qubit_num: int, n_irreps: int = 1, seed: int = 0):
    """Random real integrals with the permutational symmetries of molecular ones.

    Returns (constant, h1[m,m], chem[m,m,m,m]) with chem[a,b,c,d] = (ab|cd) 8-fold symmetric and
    both tensors zero unless the irrep labels multiply to the identity (abelian point group,
    labels in 0..n_irreps-1 combined by xor; n_irreps must be a power of two)."""
    assert qubit_num % 2 == 0
    assert n_irreps & (n_irreps - 1) == 0
    m = qubit_num // 2
    rng = np.random.default_rng(seed)
    g = rng.integers(0, n_irreps, size=m)
    h1 = rng.normal(0.0, 1.0, size=(m, m))
    h1 = 0.5 * (h1 + h1.T)
    h1 = np.where(g[:, None] == g[None, :], h1, 0.0)
    chem = 0.3 * rng.normal(0.0, 1.0, size=(m, m, m, m))
    chem = 0.5 * (chem + chem.transpose(1, 0, 2, 3))
    chem = 0.5 * (chem + chem.transpose(0, 1, 3, 2))
    chem = 0.5 * (chem + chem.transpose(2, 3, 0, 1))
    lab = g[:, None, None, None] ^ g[None, :, None, None] ^ g[None, None, :, None] ^ g[None, None, None, :]
    chem = np.where(lab == 0, chem, 0.0)
    return 0.7, h1, chem


def _ladder_masks(spin_orb: np.ndarray, n: int):
    """JW image of a_j / a+_j: 1/2 X_j (1 -/+ Z_j) Z_{<j}.  Returns (x, z0, z1) with
    a_j = 1/2 [O(x,z0) - O(x,z1)],  a+_j = 1/2 [O(x,z0) + O(x,z1)]."""
    pos = (n - 1 - spin_orb).astype(U64)
    x = U64(1) << pos
    full = U64((1 << n) - 1)
    below_incl = (U64(1) << (pos + U64(1))) - U64(1)          # bits 0..pos
    z0 = full & ~below_incl                                   # qubits 0..j-1 = bits above pos
    z1 = z0 | x
    return x, z0, z1


def _expand_product(ops, coeff, n):
    """ops: list of (spin_orb_index_array, is_dagger).  Returns flat (x, z, w) arrays of all
    2^len(ops) O(x,z) strings per input row."""
    k = coeff.shape[0]
    x = np.zeros((k, 1), dtype=U64)
    z = np.zeros((k, 1), dtype=U64)
    w = coeff.reshape(k, 1).astype(np.float64)
    for idx, dagger in ops:
        lx, z0, z1 = _ladder_masks(idx, n)
        lx = lx[:, None]
        # O(acc) O(new) = (-1)^{z_acc . x_new} O(x_acc^x_new, z_acc^z_new)
        par = (np.bitwise_count(z & lx) & 1).astype(np.float64)
        sgn = 1.0 - 2.0 * par
        nx = x ^ lx
        w0 = 0.5 * w * sgn
        w1 = w0 if dagger else -w0
        x = np.concatenate((nx, nx), axis=1)
        z = np.concatenate((z ^ z0[:, None], z ^ z1[:, None]), axis=1)
        w = np.concatenate((w0, w1), axis=1)
    return x.reshape(-1), z.reshape(-1), w.reshape(-1)


def _combine(x, z, w, tol):
    order = np.lexsort((z, x))
    x, z, w = x[order], z[order], w[order]
    new = np.ones(x.shape[0], dtype=bool)
    new[1:] = (x[1:] != x[:-1]) | (z[1:] != z[:-1])
    starts = np.flatnonzero(new)
    ws = np.add.reduceat(w, starts)
    keep = np.abs(ws) > tol
    return x[starts][keep], z[starts][keep], ws[keep]


def jordan_wigner_arrays(constant: float, h1: np.ndarray, chem: np.ndarray, tol: float = 1e-10,
                         batch: int = 1 << 20):
    """Vectorised Jordan-Wigner transform.  Returns (xy_masks u64[T], yz_masks u64[T], weights f64[T])
    sorted by (xy, yz); weights are coefficients of X^xy Z^yz (= OpenFermion coefficient * i^{#Y})."""
    m = h1.shape[0]
    n = 2 * m
    assert n <= 64
    xs, zs, ws = [np.zeros(1, U64)], [np.zeros(1, U64)], [np.array([float(constant)])]

    p, q = np.nonzero(h1)
    for s in (0, 1):
        x, z, w = _expand_product([(2 * p + s, True), (2 * q + s, False)], h1[p, q], n)
        xs.append(x); zs.append(z); ws.append(w)

    # G[p,q,r,s] = chem[p,s,q,r]; a+_{p,s} a+_{q,t} a_{r,t} a_{s,s}
    G = np.ascontiguousarray(chem.transpose(0, 2, 3, 1))
    P, Q, R, S = np.nonzero(G)
    vals = 0.5 * G[P, Q, R, S]
    acc_x, acc_z, acc_w = [], [], []
    for lo in range(0, P.shape[0], batch):
        sl = slice(lo, lo + batch)
        for s in (0, 1):
            for t in (0, 1):
                pp, qq, rr, ss = 2 * P[sl] + s, 2 * Q[sl] + t, 2 * R[sl] + t, 2 * S[sl] + s
                ok = (pp != qq) & (rr != ss)
                x, z, w = _expand_product([(pp[ok], True), (qq[ok], True), (rr[ok], False), (ss[ok], False)],
                                          vals[sl][ok], n)
                acc_x.append(x); acc_z.append(z); acc_w.append(w)
        # fold the batch early to bound memory
        x, z, w = _combine(np.concatenate(acc_x), np.concatenate(acc_z), np.concatenate(acc_w), 0.0)
        acc_x, acc_z, acc_w = [x], [z], [w]
    xs += acc_x; zs += acc_z; ws += acc_w
    return _combine(np.concatenate(xs), np.concatenate(zs), np.concatenate(ws), tol)


def synthetic_hamiltonian(qubit_num: int, n_irreps: int = 1, seed: int = 0):
    c, h1, chem = synthetic_integrals(qubit_num, n_irreps=n_irreps, seed=seed)
    return jordan_wigner_arrays(c, h1, chem)


def pauli_arrays_to_terms(xy: np.ndarray, yz: np.ndarray, w: np.ndarray, qubit_num: int) -> dict:
    """OpenFermion-style dict {((q,'X'|'Y'|'Z'),...): coeff}.  The dict coefficient multiplies the
    string written with Y's: X^x Z^z = (-i)^{#Y} * PauliString, so coeff = w * (-i)^{#Y}."""
    terms = {}
    n = qubit_num
    minus_i_pow = (1.0 + 0j, -1j, -1.0 + 0j, 1j)
    for x, z, c in zip(xy.tolist(), yz.tolist(), w.tolist()):
        key = []
        ny = 0
        both = x | z
        for qb in range(n):
            bit = 1 << (n - 1 - qb)
            if both & bit:
                if (x & bit) and (z & bit):
                    key.append((qb, 'Y')); ny += 1
                elif x & bit:
                    key.append((qb, 'X'))
                else:
                    key.append((qb, 'Z'))
        terms[tuple(key)] = c * minus_i_pow[ny & 3]
    return terms


def random_physical_samples(qubit_num: int, alpha_num: int, beta_num: int, count: int, seed: int = 1) -> np.ndarray:
    """`count` distinct configurations with alpha_num set bits on even positions (mask 0x5555...,
    pauli_observable.py:553) and beta_num on odd positions, sorted ascending."""
    assert qubit_num % 2 == 0
    m = qubit_num // 2
    rng = np.random.default_rng(seed)
    out = np.zeros(0, dtype=U64)
    while out.shape[0] < count:
        need = int((count - out.shape[0]) * 1.1) + 16
        ka = np.argsort(rng.random((need, m)), axis=1)[:, :alpha_num].astype(U64)
        kb = np.argsort(rng.random((need, m)), axis=1)[:, :beta_num].astype(U64)
        a = np.bitwise_or.reduce(U64(1) << (U64(2) * ka), axis=1) if alpha_num else np.zeros(need, U64)
        b = np.bitwise_or.reduce(U64(1) << (U64(2) * kb + U64(1)), axis=1) if beta_num else np.zeros(need, U64)
        out = np.unique(np.concatenate((out, a | b)))
        total = _sector_size(m, alpha_num, beta_num)
        if total <= count:
            break
    if out.shape[0] > count:
        out = np.sort(rng.permutation(out)[:count])
    return out


def clustered_physical_samples(qubit_num: int, alpha_num: int, beta_num: int, count: int, seed: int = 1,
                               mean_rank: float = 3.0) -> np.ndarray:
    """`count` distinct physical configurations concentrated around the Hartree-Fock determinant (lowest orbitals
    occupied = highest bits set, create_masker.py:29): each one is HF with k ~ 1 + Poisson(mean_rank - 1) random
    single excitations (occupied -> virtual within one spin sector).  This is what a VMC sample set looks like - many
    samples share their alpha (or beta) string and connected configurations are often sampled too - as opposed to
    `random_physical_samples`, whose uniform draws almost never connect to each other at large qubit counts."""
    assert qubit_num % 2 == 0
    m = qubit_num // 2
    rng = np.random.default_rng(seed)
    hf_a = np.zeros(m, bool); hf_a[m - alpha_num:] = True
    hf_b = np.zeros(m, bool); hf_b[m - beta_num:] = True
    total = _sector_size(m, alpha_num, beta_num)
    count = min(count, total)
    out = np.zeros(0, dtype=U64)
    weights = U64(1) << (U64(2) * np.arange(m, dtype=U64))
    while out.shape[0] < count:
        need = int((count - out.shape[0]) * 1.5) + 64
        occ = np.stack((np.tile(hf_a, (need, 1)), np.tile(hf_b, (need, 1))), axis=1)  # [need, 2, m]
        k = 1 + rng.poisson(max(mean_rank - 1.0, 0.0), size=need)
        rows = np.arange(need)
        for step in range(int(k.max())):
            act = k > step
            spin = rng.integers(0, 2, size=need)
            cur = occ[rows, spin]                                      # [need, m]
            r = rng.random((need, m))
            src = np.argmax(np.where(cur, r, -1.0), axis=1)            # a random occupied orbital
            dst = np.argmax(np.where(~cur, r, -1.0), axis=1)           # a random empty orbital
            ok = act & cur.any(axis=1) & (~cur).any(axis=1)
            occ[rows[ok], spin[ok], src[ok]] = False
            occ[rows[ok], spin[ok], dst[ok]] = True
        a = (occ[:, 0, :] * weights).sum(axis=1, dtype=U64)
        b = (occ[:, 1, :] * (weights << U64(1))).sum(axis=1, dtype=U64)
        out = np.unique(np.concatenate((out, a | b)))
        mean_rank += 0.25  # widen the cloud if the shell is exhausted
    if out.shape[0] > count:
        out = np.sort(rng.permutation(out)[:count])
    return out


def _sector_size(m, na, nb):
    from math import comb
    return comb(m, na) * comb(m, nb)


def all_physical_samples(qubit_num: int, alpha_num: int, beta_num: int) -> np.ndarray:
    """The whole (N_alpha, N_beta) sector, ascending.  Small qubit counts only."""
    from itertools import combinations
    m = qubit_num // 2
    a = [sum(1 << (2 * k) for k in c) for c in combinations(range(m), alpha_num)]
    b = [sum(1 << (2 * k + 1) for k in c) for c in combinations(range(m), beta_num)]
    out = (np.array(a, dtype=U64)[:, None] | np.array(b, dtype=U64)[None, :]).reshape(-1)
    return np.sort(out)


def random_amplitudes(count: int, seed: int = 2) -> np.ndarray:
    """log|psi| ~ N(0, 2^2), phase ~ U(-pi, pi), normalised."""
    rng = np.random.default_rng(seed)
    log_abs = rng.normal(0.0, 2.0, size=count)
    phase = rng.uniform(-np.pi, np.pi, size=count)
    amps = np.exp(log_abs + 1j * phase)
    return (amps / np.sqrt(np.sum(np.abs(amps) ** 2))).astype(np.complex128)
