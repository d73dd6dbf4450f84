"""Exciton-number sector geometry of the ket x bra blocks the CUDA engine stores.

Every operation of the reference's step (qevolve.py:74-348) conserves the number of excited chromophores:
the system block is Z_i and X_iX_j + Y_iY_j, and pseudomodes / collisions touch system qubits only through Z_i.
So sigma splits into blocks (KA excitons on the ket side, KB on the bra side) that never mix, and only the
blocks a calculation needs are stored.  Index convention (shared with include/qsx.h):

    a = cfg_index << n_bits | bits        cfg_index enumerates the N-bit occupation masks with KA bits set,
                                          in increasing order; ``bits`` = pseudomode (+ancilla) qubits.
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import lru_cache
from itertools import combinations

import numpy as np


@lru_cache(maxsize=None)
def sector_configs(n_sites: int, k: int) -> tuple[int, ...]:
    """Occupation masks (bit i = chromophore i excited) with exactly k bits set, increasing."""
    return tuple(sorted(sum(1 << i for i in c) for c in combinations(range(n_sites), k)))


@lru_cache(maxsize=None)
def sector_rank(n_sites: int, k: int) -> np.ndarray:
    """mask -> position in ``sector_configs`` or -1."""
    rank = np.full(1 << n_sites, -1, dtype=np.int32)
    for idx, mask in enumerate(sector_configs(n_sites, k)):
        rank[mask] = idx
    return rank


@dataclass(frozen=True)
class Block:
    """Geometry of one (KA, KB) block."""
    n_sites: int
    n_bits: int
    ka: int
    kb: int

    @property
    def cfg_a(self):
        return sector_configs(self.n_sites, self.ka)

    @property
    def cfg_b(self):
        return sector_configs(self.n_sites, self.kb)

    @property
    def n_a(self):
        return len(self.cfg_a)

    @property
    def n_b(self):
        return len(self.cfg_b)

    @property
    def dim_a(self):
        return self.n_a << self.n_bits

    @property
    def dim_b(self):
        return self.n_b << self.n_bits

    @property
    def size(self):
        return self.dim_a * self.dim_b

    def valid(self) -> bool:
        return 0 <= self.ka <= self.n_sites and 0 <= self.kb <= self.n_sites

    # ---- read-out tables (CSR: ptr, idx, complex weights) -----------------------------------
    def config_probability_observables(self):
        """One observable per ket configuration: sum over bits of the diagonal (needs ka == kb)."""
        assert self.ka == self.kb
        nbits = 1 << self.n_bits
        ptr, idx = [0], []
        for c in range(self.n_a):
            for m in range(nbits):
                a = (c << self.n_bits) | m
                idx.append(a * self.dim_b + a)
            ptr.append(len(idx))
        return np.array(ptr, np.int32), np.array(idx, np.int32), np.ones(len(idx), complex)

    def emission_observable(self, mu):
        """Tr[(sum_s mu_s X_s) sigma]: elements with equal bits and occupation masks differing in site s."""
        rank_a = sector_rank(self.n_sites, self.ka)
        nbits = 1 << self.n_bits
        idx, w = [], []
        for cb, mask_b in enumerate(self.cfg_b):
            for s in range(self.n_sites):
                ca = rank_a[mask_b ^ (1 << s)]
                if ca < 0:
                    continue
                for m in range(nbits):
                    a, b = (ca << self.n_bits) | m, (cb << self.n_bits) | m
                    idx.append(a * self.dim_b + b)
                    w.append(mu[s])
        return np.array([0, len(idx)], np.int32), np.array(idx, np.int32), np.array(w, complex)


def split_state_by_sector(n_sites: int, amplitudes: np.ndarray) -> dict[int, np.ndarray]:
    """2**N amplitudes -> {K: amplitudes over sector_configs(N, K)} for the sectors that are populated."""
    out = {}
    for k in range(n_sites + 1):
        cfgs = np.array(sector_configs(n_sites, k), dtype=np.int64)
        part = amplitudes[cfgs]
        if np.any(part != 0):
            out[k] = part
    return out


def population_observables(block: Block):
    """One observable per chromophore: P_i = sum of the diagonal elements whose configuration has site i excited."""
    assert block.ka == block.kb
    nbits = 1 << block.n_bits
    ptr, idx = [0], []
    for site in range(block.n_sites):
        for c, mask in enumerate(block.cfg_a):
            if (mask >> site) & 1:
                for m in range(nbits):
                    a = (c << block.n_bits) | m
                    idx.append(a * block.dim_b + a)
        ptr.append(len(idx))
    return np.array(ptr, np.int32), np.array(idx, np.int32), np.ones(len(idx), complex)
