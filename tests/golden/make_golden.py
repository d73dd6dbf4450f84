"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) over oracle/refshim.

Run in the build container only (the reference is not on the GPU box):

    python tests/golden/make_golden.py

Each fixture stores the inputs of one reference call (as JSON) and its exact output: outcome
probabilities of ``qevolve`` circuits, the complex ``qspectroscopy``/``clspectroscopy`` signal, the
``clevolve`` populations, the dense Hamiltonians, and the Pauli dictionaries of ``__mode_operators``.
``shots`` only sizes the (ignored) sampling; the shim executes exactly (shots -> infinity limit).
"""
import json
import os
import sys
import time
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(REPO, 'oracle', 'refshim'), '/root/reference', REPO]
warnings.simplefilter('ignore')

import numpy as np  # noqa: E402
from scipy.linalg import toeplitz  # noqa: E402

import qsextra  # noqa: E402
assert qsextra.__file__.startswith('/root/reference'), qsextra.__file__
from qsextra import ExcitonicSystem, ChromophoreSystem  # noqa: E402
from qsextra.qcomo import qevolve, clevolve  # noqa: E402
from qsextra.spectroscopy import FeynmanDiagram, Spectroscopy, qspectroscopy, clspectroscopy  # noqa: E402
from qsextra.spectroscopy.postprocessing import postprocessing  # noqa: E402

HBAR = 0.658212


def build(sysdef):
    kw = dict(energies=sysdef['energies'], couplings=np.array(sysdef['couplings'], dtype=float)
              if not np.isscalar(sysdef['couplings']) else sysdef['couplings'],
              dipole_moments=sysdef.get('dipole_moments'))
    if 'frequencies_pseudomode' in sysdef:
        s = ChromophoreSystem(**kw, frequencies_pseudomode=sysdef['frequencies_pseudomode'],
                              levels_pseudomode=sysdef['levels_pseudomode'], couplings_ep=sysdef['couplings_ep'])
    else:
        s = ExcitonicSystem(**kw)
    if 'state' in sysdef:
        st = sysdef['state']
        if st[0] in ('state', 'delocalized excitation'):
            s.set_state(st[0], [complex(*c) if isinstance(c, list) else c for c in st[1]])
        else:
            s.set_state(st[0], st[1])
    return s


def save(name, meta, **arrays):
    np.savez_compressed(os.path.join(HERE, name + '.npz'), meta=json.dumps(meta), **arrays)
    print('wrote', name, {k: np.shape(v) for k, v in arrays.items()})


DIMER = dict(energies=[1.55, 1.46], couplings=-0.01, dipole_moments=[1., 1.])
DT_FS = 0.1 / HBAR
GAMMA = 59.08e-3


def qevolve_cases():
    ex1 = dict(energies=[1., 2., 3., 4.], couplings=toeplitz([0., 1., 0., 1.]).tolist(), dipole_moments=[1.] * 4)
    ex2 = dict(energies=[1., 2.], couplings=1., dipole_moments=[1., 1.],
               frequencies_pseudomode=0., levels_pseudomode=2, couplings_ep=1.)
    cases = {
        'ex1_unitary': (dict(ex1, state=['localized excitation', 0]),
                        dict(time=np.arange(0, 2.01, 0.1).tolist(), dt=0.01)),
        'ex1_dephasing': (dict(ex1, state=['delocalized excitation', [1 / np.sqrt(2), 1 / np.sqrt(2), 0, 0]]),
                          dict(time=np.arange(0, 2.01, 0.1).tolist(), dt=0.01, coll_rates=0.1)),
        'ex2_pseudomode': (dict(ex2, state=['localized excitation', 0]),
                           dict(time=np.arange(0, 5.01, 0.5).tolist(), dt=0.01, coll_rates=2.)),
        'c1_dimer_markov': (dict(DIMER, state=['localized excitation', 0]),
                            dict(time=(np.arange(0, 201, 20) * DT_FS).tolist(), dt=DT_FS, coll_rates=GAMMA / 4)),
        'dimer_two_modes': (dict(energies=[1.55, 1.46], couplings=-0.01, dipole_moments=[1., 0.8],
                                 frequencies_pseudomode=[0.02, 0.11], levels_pseudomode=[2, 2],
                                 couplings_ep=[0.03, 0.05], state=['localized excitation', 1]),
                            dict(time=(np.arange(0, 101, 25) * DT_FS).tolist(), dt=DT_FS, coll_rates=[0.2, 0.07])),
        'general_state_trotter2': (dict(energies=[1.0, 1.3, 0.7], couplings=[[0, .2, .05], [.2, 0, .3], [.05, .3, 0]],
                                        dipole_moments=[1., 1., 1.],
                                        state=['state', [[.5, 0], [.1, .3], [0, -.4], [.2, 0], [.3, .1], [0, .2], [-.1, 0], [.4, -.2]]]),
                                   dict(time=np.arange(0, 1.01, 0.2).tolist(), dt=0.1, Trotter_number=2, coll_rates=0.3)),
        'no_dt_given': (dict(DIMER, state=['localized excitation', 0]),
                        dict(time=np.arange(0, 3.01, 0.5).tolist(), Trotter_number=5)),
        'suzuki2_trimer': (dict(energies=[1.0, 1.3, 0.7], couplings=[[0, .2, .05], [.2, 0, .3], [.05, .3, 0]],
                                dipole_moments=[1., 1., 1.], state=['localized excitation', 2]),
                           dict(time=np.arange(0, 2.01, 0.5).tolist(), dt=0.25, Trotter_order=2, coll_rates=0.05)),
        'suzuki4_dimer_modes': (dict(energies=[1., 2.], couplings=1., dipole_moments=[1., 1.],
                                     frequencies_pseudomode=0.3, levels_pseudomode=2, couplings_ep=0.7,
                                     state=['localized excitation', 0]),
                                dict(time=np.arange(0, 1.01, 0.25).tolist(), dt=0.25, Trotter_order=4, coll_rates=1.5)),
        'mode_d3': (dict(energies=[1., 1.4], couplings=0.3, dipole_moments=[1., 1.],
                         frequencies_pseudomode=0.2, levels_pseudomode=3, couplings_ep=0.5,
                         state=['localized excitation', 0]),
                    dict(time=np.arange(0, 1.01, 0.2).tolist(), dt=0.05, coll_rates=0.8)),
        'mode_d4_no_collisions': (dict(energies=[1.2], couplings=[[0.]], dipole_moments=[1.],
                                       frequencies_pseudomode=0.4, levels_pseudomode=4, couplings_ep=0.6,
                                       state=['localized excitation', 0]),
                                  dict(time=np.arange(0, 1.01, 0.25).tolist(), dt=0.05)),
        'mode_d4_collisions': (dict(energies=[1.2, 0.9], couplings=0.2, dipole_moments=[1., 1.],
                                    frequencies_pseudomode=0.4, levels_pseudomode=4, couplings_ep=0.6,
                                    state=['delocalized excitation', [0.6, 0.8]]),
                               dict(time=np.arange(0, 0.51, 0.1).tolist(), dt=0.05, coll_rates=0.9)),
        'ground_state': (dict(DIMER, state=['ground', 0]),
                         dict(time=[0., DT_FS * 10], dt=DT_FS, coll_rates=GAMMA / 4)),
    }
    for name, (sysdef, kw) in cases.items():
        t0 = time.time()
        system = build(sysdef)
        res = qevolve(system, kw['time'] if len(kw['time']) > 1 else kw['time'][0],
                      shots=16, verbose=False, **{k: v for k, v in kw.items() if k != 'time'})
        dist = np.array(res.exact_probabilities)
        save('qevolve_' + name, dict(system=sysdef, kwargs=kw), distribution=dist)
        print('   %.1fs' % (time.time() - t0))


def spectroscopy_cases():
    t7 = (np.arange(0, 70, 10) / HBAR).tolist()
    t3 = (np.arange(0, 30, 10) / HBAR).tolist()
    t2s = (np.arange(0, 20, 10) / HBAR).tolist()
    dimer_modes = dict(DIMER, frequencies_pseudomode=0.02, levels_pseudomode=2,
                       couplings_ep=float(np.sqrt(GAMMA * 0.1 / 2)))
    trimer = dict(energies=[1.55, 1.46, 1.50], couplings=[[0, -.01, .004], [-.01, 0, -.008], [.004, -.008, 0]],
                  dipole_moments=[1., 0.7, -0.4])
    mk = dict(dt=DT_FS, coll_rates=GAMMA / 4)
    cases = {
        'abs_dimer_markov': (DIMER, 'a', t7, mk),
        'abs_dimer_unitary': (DIMER, 'a', t7, dict(dt=DT_FS)),
        'abs_dimer_modes': (dimer_modes, 'a', t3, dict(dt=DT_FS, coll_rates=[0.2])),
        'gsb_dimer_markov': (DIMER, 'gsb', [t3, [0.], t3], mk),
        'se_dimer_markov': (DIMER, 'se', [t3, [0.], t3], mk),
        'esa_dimer_markov': (DIMER, 'esa', [t3, [0.], t3], mk),
        'se_dimer_t2': (DIMER, 'se', [t2s, [0., 30 / HBAR], t2s], mk),
        'esa_dimer_t2': (DIMER, 'esa', [t2s, [0., 30 / HBAR], t2s], mk),
        'gsb_dimer_modes': (dimer_modes, 'gsb', [t2s, [5 / HBAR], t2s], dict(dt=DT_FS, coll_rates=[0.2])),
        'esa_dimer_modes': (dimer_modes, 'esa', [t2s, [5 / HBAR], t2s], dict(dt=DT_FS, coll_rates=[0.2])),
        'se_trimer_markov': (trimer, 'se', [[0., 2.], [1.], [0., 3.]], dict(dt=0.5, coll_rates=0.02)),
        'esa_trimer_trotter': (trimer, 'esa', [[1.], [0., 2.], [2.]], dict(dt=1.0, Trotter_number=2, coll_rates=0.02)),
        # non-rephasing pathways via hand-set sequences on the base class
        # unequal dipoles: the emission / dipole weights of the two sites are distinguishable
        'esa_dimer_unequal_mu': (dict(DIMER, dipole_moments=[1., 0.8]), 'esa', [t2s, [0., 30 / HBAR], t2s], mk),
        'gsb_dimer_modes_unequal_mu': (dict(dimer_modes, dipole_moments=[0.7, 1.2]), 'gsb', [t2s, [5 / HBAR], t2s],
                                       dict(dt=DT_FS, coll_rates=[0.2])),
        'esa_dimer_modes_unequal_mu': (dict(dimer_modes, dipole_moments=[0.7, 1.2]), 'esa', [t2s, [5 / HBAR], t2s],
                                       dict(dt=DT_FS, coll_rates=[0.2])),
        'custom_kkkk_dimer': (DIMER, ('kkkk', 'ioio'), [t2s, [0.], t2s], mk),
        'custom_kbbk_dimer': (DIMER, ('kbbk', 'iioo'), [t2s, [0.], t2s], mk),
        'custom_kbkk_dimer': (DIMER, ('kbkk', 'iiio'), [t2s, [0.], t2s], mk),
        # dt not given: every scalar delay becomes its own step size (delay / Trotter_number), qevolve.py:373-379
        'abs_dimer_no_dt': (DIMER, 'a', [0.7, 1.9, 0.4], dict(Trotter_number=3, coll_rates=GAMMA / 4)),
        'esa_dimer_modes_no_dt': (dimer_modes, 'esa', [[0.5, 1.1], [0.8], [0.3, 0.9]], dict(Trotter_number=2, coll_rates=[0.2])),
    }
    only = [a for a in sys.argv[1:] if a.startswith('only=')]
    for name, (sysdef, label, delays, kw) in cases.items():
        if only and name not in only[0][5:].split(','):
            continue
        t0 = time.time()
        system = build(sysdef)
        if isinstance(label, tuple):
            spec = Spectroscopy(delays)
            spec.side_seq, spec.direction_seq = label
        else:
            spec = FeynmanDiagram(label, delays)
        signal = qspectroscopy(system, spec, shots=16, verbose=False, **dict(kw))
        arrays = dict(signal=signal)
        if 'unitary' not in name and 'trotter' not in name:
            rates = kw.get('coll_rates')
            arrays['signal_cl'] = clspectroscopy(system, spec, verbose=False, rates=rates)
        save('qspectroscopy_' + name, dict(system=sysdef, label=label, delay_time=delays, kwargs=kw), **arrays)
        print('   %.1fs' % (time.time() - t0))


def classical_cases():
    ex1 = dict(energies=[1., 2., 3., 4.], couplings=toeplitz([0., 1., 0., 1.]).tolist(), dipole_moments=[1.] * 4,
               state=['localized excitation', 0])
    ex2 = dict(energies=[1., 2.], couplings=1., dipole_moments=[1., 1.],
               frequencies_pseudomode=0., levels_pseudomode=2, couplings_ep=1., state=['localized excitation', 0])
    for name, sysdef, time_list, rates in [('ex1_dephasing', ex1, np.arange(0, 2.01, 0.1), 0.1),
                                           ('ex1_unitary', ex1, np.arange(0, 2.01, 0.1), None),
                                           ('ex2_pseudomode', ex2, np.arange(0, 5.01, 0.5), [2.])]:
        system = build(sysdef)
        res = clevolve(system, time_list, rates=rates, verbose=False)
        save('clevolve_' + name, dict(system=sysdef, time=time_list.tolist(), rates=rates),
             populations=np.array(res.expect).T)


def structure_cases():
    ex1 = build(dict(energies=[1., 2., 3., 4.], couplings=toeplitz([0., 1., 0., 1.]).tolist(), dipole_moments=[1.] * 4))
    ex2 = build(dict(energies=[1., 2.], couplings=1., dipole_moments=[1., 1.],
                     frequencies_pseudomode=0., levels_pseudomode=2, couplings_ep=1.))
    three = build(dict(energies=[1., 2.], couplings=0.5, dipole_moments=[1., 1.],
                       frequencies_pseudomode=[0.3, 0.7], levels_pseudomode=[3, 2], couplings_ep=[0.1, 0.2]))
    save('hamiltonians', {},
         H4=ex1.get_e_Hamiltonian().full(), H2_e=ex2.get_e_Hamiltonian().full(),
         H2_p=ex2.get_p_Hamiltonian().full(), H2_ep=ex2.get_ep_Hamiltonian().full(),
         H2_global=ex2.get_global_Hamiltonian().full(), H32_global=three.get_global_Hamiltonian().full())
    qmod = sys.modules['qsextra.qcomo.evolution.qevolve']
    mode_ops = qmod.__dict__['__mode_operators']
    out = {}
    for d in (2, 3, 4, 5, 8, 16):
        s = build(dict(energies=[1.], couplings=[[0.]], dipole_moments=[1.],
                       frequencies_pseudomode=1., levels_pseudomode=d, couplings_ep=1.))
        dicts = mode_ops(s, 0)
        out[str(d)] = [[[k, [np.real(v), np.imag(v)]] for k, v in dd.items()] for dd in dicts]
    save('mode_operators', out)
    # post-processing of a synthetic signal (reference postprocessing.py:99-122)
    rng = np.random.default_rng(7)
    t = (np.arange(0, 50, 10) / HBAR)
    lin = rng.normal(size=5) + 1j * rng.normal(size=5)
    sig = rng.normal(size=(5, 2, 5)) + 1j * rng.normal(size=(5, 2, 5))
    w1, s1 = postprocessing(FeynmanDiagram('a', t), lin, RF_freq=1.505, damping_rate=0.01, pad_extension=3)
    w2, s2 = postprocessing(FeynmanDiagram('gsb', [t, [0., 1.], t]), sig, RF_freq=1.505, damping_rate=0.01,
                            pad_extension=3, T2_index=1)
    save('postprocessing', dict(t=t.tolist()), lin=lin, sig=sig, omega_lin=w1, spec_lin=s1,
         omega1=w2[0], omega3=w2[1], spec_2d=s2)


def heom_benchmark():
    """The only numeric artefact the reference ships for the master-equation path: the HEOM curve rho_11(t) of
    ``Examples/qcomo - Examples/Benchmark Example 3/data.mat`` (401 points, t = 0 ... 20), which the notebook
    'Example 3 - More is better' overlays on ``clevolve`` with 1 x 16-, 2 x 4- and 4 x 2-level pseudomodes per site
    (cells 8-26: N = 2, eps = 0, J = 1, Gamma = 20, Omega = 0.1, g = sqrt(Gamma Omega / (2 W)), rates 2 Omega)."""
    from scipy.io import loadmat
    m = loadmat('/root/reference/Examples/qcomo - Examples/Benchmark Example 3/data.mat')
    save('heom_example3', dict(source='Examples/qcomo - Examples/Benchmark Example 3/data.mat', N=2, energies=[0., 0.],
                               coupling=1.0, Gamma=20.0, Omega=0.1),
         t=m['x'][0], rho11=m['y'][0])


if __name__ == '__main__':
    which = [a for a in sys.argv[1:] if not a.startswith('only=')] or ['structure', 'qevolve', 'classical', 'spectroscopy', 'heom']
    if 'heom' in which:
        heom_benchmark()
    if 'structure' in which:
        structure_cases()
    if 'qevolve' in which:
        qevolve_cases()
    if 'classical' in which:
        classical_cases()
    if 'spectroscopy' in which:
        spectroscopy_cases()
