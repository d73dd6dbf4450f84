"""Shared test helpers: golden fixtures -> product objects."""
import glob
import json
import os

import numpy as np

from qsextra_b200 import ChromophoreSystem, ExcitonicSystem
from qsextra_b200.spectroscopy import FeynmanDiagram, Spectroscopy

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden_names(prefix):
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, prefix + '*.npz')))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    return json.loads(str(z['meta'])), {k: z[k] for k in z.files if k != 'meta'}


def build_system(sd):
    couplings = sd['couplings'] if np.isscalar(sd['couplings']) else np.array(sd['couplings'], dtype=float)
    kw = dict(energies=sd['energies'], couplings=couplings, dipole_moments=sd.get('dipole_moments'))
    if 'frequencies_pseudomode' in sd:
        s = ChromophoreSystem(**kw, frequencies_pseudomode=sd['frequencies_pseudomode'],
                              levels_pseudomode=sd['levels_pseudomode'], couplings_ep=sd['couplings_ep'])
    else:
        s = ExcitonicSystem(**kw)
    if 'state' in sd:
        kind, value = sd['state']
        if kind in ('state', 'delocalized excitation'):
            value = [complex(*c) if isinstance(c, list) else c for c in value]
        s.set_state(kind, value)
    return s


def build_spectroscopy(label, delays):
    if isinstance(label, (list, tuple)):
        spec = Spectroscopy(delays)
        spec.side_seq, spec.direction_seq = label
        return spec
    return FeynmanDiagram(label, delays)


def time_arg(t):
    return t if len(t) > 1 else t[0]
