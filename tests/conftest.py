import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'reference: needs the read-only reference tree at /root/reference')


def load_golden(name):
    return dict(np.load(os.path.join(GOLD, name)))


@pytest.fixture(scope='session')
def treynor():
    return load_golden('treynor_inputs.npz')


@pytest.fixture(scope='session')
def treynor_kw():
    return load_golden('treynor_kw_standalone.npz')


@pytest.fixture(scope='session')
def treynor_dw():
    """The same storm on the reference's DIFFUSIVE-wave component, dt = 2 s, 3600 steps
    (oracle/gen_golden.py gen_treynor_dw)."""
    return load_golden('treynor_dw_standalone.npz')


@pytest.fixture(scope='session')
def synth_channels():
    return load_golden('synth_channels.npz')


@pytest.fixture(scope='session')
def d8_synth():
    return load_golden('d8_synth.npz')


@pytest.fixture(scope='session')
def sources_gold():
    return load_golden('sources.npz')


def rel_err(a, b):
    """max |a-b| / max(|b|) over the array (0 when both are all-zero)."""
    a = np.asarray(a, dtype='float64')
    b = np.asarray(b, dtype='float64')
    scale = np.abs(b).max()
    if scale == 0:
        return float(np.abs(a).max())
    return float(np.abs(a - b).max() / scale)


def cell_rel_err(a, b, floor_frac=1e-12):
    """Per-cell relative error  max_i |a_i - b_i| / max(|b_i|, floor)  with the absolute
    floor = floor_frac * max|b|.  rel_err() above normalises by the grid maximum and is
    blind to large relative errors in small-valued cells (most of a grid early in a
    storm); this one is not.  The floor is needed because cells whose volume the
    reference clamps to exactly 0 (np.maximum(vol, 0), channels_base.py:2145) may hold
    a last-bit residue here: below 1e-12 of the grid maximum a difference counts as
    absolute."""
    a = np.asarray(a, dtype='float64')
    b = np.asarray(b, dtype='float64')
    scale = np.abs(b).max()
    if scale == 0:
        return float(np.abs(a).max())
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor_frac * scale)).max())


@pytest.fixture(scope='session')
def flood_gold():
    return load_golden('flood_channels.npz')


@pytest.fixture(scope='session')
def met_gold():
    return load_golden('met_energy.npz')


@pytest.fixture(scope='session')
def attenuate_gold():
    return load_golden('attenuate_channels.npz')


@pytest.fixture(scope='session')
def geo_gold():
    return load_golden('geo_channels.npz')
