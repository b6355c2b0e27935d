"""Exact diagonalisation / known-answer tooling.  TEST INFRASTRUCTURE ONLY (same rule as rnnwf_oracle.py: only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import it).

Sparse-Lanczos restatement (SciPy `eigsh`) of the dense helpers the reference keeps in its notebooks
(`IsingMatrixElements` / `ED_1DTFIM`, Tutorials/1DTFIM/Tutorial_1DTFIM.ipynb#cell6; `J1J2MatrixElements` / `ED_j1j2`,
Tutorials/J1J2/Tutorial_1DJ1J2.ipynb#cell6), extended to the open 2-D lattice the 2-D apps train on
(2DTFIM_1DRNN/run_2dTFIM.py:10 uses 4x4, Bx=3) and to exact values of the observables in
rnnwavefunctions_b200/observables.py.  Pinned in tests/test_oracle.py against the energies recorded in the notebooks
(tests/golden/known_answers.npz) and the free-fermion formula.

Conventions (those of the reference's local-energy functions): a configuration is a 0/1 vector in site order, site 0 is the
MOST significant bit of the basis index, sigma^z = 2 v - 1;
    H_TFIM = - sum_<ij> Jz_ij sz_i sz_j - Bx sum_i sx_i                  (1DTFIM/TrainingRNN_1DTFIM.py:31-38,74)
    H_J1J2 = sum_i J1_i S_i.S_{i+1} + J2_i S_i.S_{i+2}, open chain; Marshall rotation flips the sign of the J1 exchange
                                                                          (J1J2/TrainingRNN_J1J2.py:32-92)
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def _bits(N):
    idx = np.arange(1 << N, dtype=np.int64)
    return idx, ((idx[:, None] >> np.arange(N - 1, -1, -1)[None, :]) & 1).astype(np.int8)


def _lowest(H, k=1):
    dim = H.shape[0]
    if dim <= 512:
        w, v = np.linalg.eigh(H.toarray())
        return w[0], v[:, 0]
    w, v = spla.eigsh(H.tocsr(), k=k, which="SA", tol=1e-12, ncv=max(20, 4 * k))
    o = np.argsort(w)
    return w[o[0]], v[:, o[0]]


def tfim_hamiltonian(Jz, Bx):
    """Sparse H of the open chain (Jz [N], bond i couples sites i, i+1) or the open lattice (Jz [Nx, Ny]: bond (i,j)-(i+1,j)
    weighs Jz[i,j], bond (i,j)-(i,j+1) weighs Jz[i,j], as 2DTFIM_2DRNN/Training2DRNN_2DTFIM.py:33-49; site order x-major)."""
    Jz = np.asarray(Jz, dtype=np.float64)
    if Jz.ndim == 1:
        N = Jz.shape[0]
        bonds = [(i, i + 1, Jz[i]) for i in range(N - 1)]
    else:
        Nx, Ny = Jz.shape
        N = Nx * Ny
        bonds = [(i * Ny + j, (i + 1) * Ny + j, Jz[i, j]) for i in range(Nx - 1) for j in range(Ny)]
        bonds += [(i * Ny + j, i * Ny + j + 1, Jz[i, j]) for i in range(Nx) for j in range(Ny - 1)]
    idx, b = _bits(N)
    s = 2.0 * b - 1.0
    diag = np.zeros(1 << N)
    for i, j, J in bonds:
        diag -= J * s[:, i] * s[:, j]
    H = sp.diags(diag).tocsr()
    if Bx != 0:
        rows = np.concatenate([idx] * N)
        cols = np.concatenate([idx ^ (1 << (N - 1 - i)) for i in range(N)])
        H = H + sp.csr_matrix((np.full(rows.shape, -float(Bx)), (rows, cols)), shape=H.shape)
    return H


def tfim_ground_state(Jz, Bx):
    """-> (E0, psi [2^N], real, normalised, non-negative: the TFIM ground state is sign-free for Bx > 0)."""
    e, v = _lowest(tfim_hamiltonian(Jz, Bx))
    if v.sum() < 0:
        v = -v
    return float(e), v


def j1j2_hamiltonian(N, J1, J2, marshall_sign=False):
    """Sparse H of the open J1-J2 chain in the full 2^N space (spin-1/2 operators S = sigma/2)."""
    J1 = np.broadcast_to(np.asarray(J1, dtype=np.float64), (N,))
    J2 = np.broadcast_to(np.asarray(J2, dtype=np.float64), (N,))
    idx, b = _bits(N)
    s = b - 0.5
    diag = np.zeros(1 << N)
    rows, cols, vals = [], [], []
    for dist, J, sign in ((1, J1, -1.0 if marshall_sign else 1.0), (2, J2, 1.0)):
        for i in range(N - dist):
            j = i + dist
            diag += J[i] * s[:, i] * s[:, j]
            anti = b[:, i] != b[:, j]
            flip = (1 << (N - 1 - i)) | (1 << (N - 1 - j))
            rows.append(idx[anti])
            cols.append(idx[anti] ^ flip)
            vals.append(np.full(int(anti.sum()), sign * 0.5 * J[i]))
    H = sp.diags(diag).tocsr()
    H = H + sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=H.shape)
    return H


def j1j2_ground_state(N, J1=1.0, J2=0.0, marshall_sign=False, zero_magnetisation=True):
    """-> (E0, psi).  With `zero_magnetisation` the search is restricted to sum(v) = N/2 (the sector the cRNN samples,
    J1J2/ComplexRNNwavefunction.py:85-93); psi is returned in the full basis."""
    H = j1j2_hamiltonian(N, J1, J2, marshall_sign)
    if not zero_magnetisation:
        return _lowest(H)
    _, b = _bits(N)
    keep = np.nonzero(b.sum(1) == N // 2)[0]
    e, v = _lowest(H[keep][:, keep])
    psi = np.zeros(1 << N)
    psi[keep] = v
    return float(e), psi


# ---- exact observables of a state vector (checks of rnnwavefunctions_b200/observables.py) ----------------------------
def sz_moments(psi):
    N = int(np.log2(psi.shape[0]))
    _, b = _bits(N)
    s = 2.0 * b - 1.0
    p = np.abs(psi) ** 2
    return p @ s, (s * p[:, None]).T @ s


def sigma_x(psi):
    N = int(np.log2(psi.shape[0]))
    idx = np.arange(psi.shape[0])
    return np.array([np.real(np.vdot(psi, psi[idx ^ (1 << (N - 1 - i))])) for i in range(N)])


def renyi2(psi, n_A):
    """S_2 of the first n_A sites."""
    N = int(np.log2(psi.shape[0]))
    M = psi.reshape(1 << n_A, 1 << (N - n_A))
    rho = M @ M.conj().T
    return float(-np.log(np.real(np.trace(rho @ rho))))
