"""Live pin against the real third-party routine, when it is importable (it is not in the build image).

The oracle is "parity unpinned" because powerbox is absent here.  On a machine that has powerbox these tests pin
the oracle (edges, counts, power) and the edge-rule detection against ``powerbox.get_power`` itself.
"""
import numpy as np
import pytest

import py21cmmc_b200  # noqa: F401
from py21cmmc_b200 import binning
from oracle import ps_oracle as O

powerbox = pytest.importorskip("powerbox")


@pytest.mark.parametrize("shape,length,log_bins", [((16, 16, 24), (16.0, 16.0, 24.0), True),
                                                   ((12, 12, 12), (9.0, 9.0, 9.0), False),
                                                   ((10, 10, 15), (10.0, 10.0, 5.0), True)])
def test_oracle_matches_installed_powerbox(shape, length, log_bins):
    from powerbox.tools import get_power

    rule = binning.detect_edge_rule()
    assert rule in binning.EDGE_RULES
    box = (np.random.default_rng(3).normal(size=shape) * 10 + 20).astype(np.float32)
    w = O.k_weights_lightcone(shape)
    res = get_power(box, boxlength=list(length), bins=7, bin_ave=False, get_variance=False, log_bins=log_bins,
                    k_weights=w)
    p_o, e_o = O.get_power(box, length, bins=7, log_bins=log_bins, k_weights=w, edge_rule=rule)
    assert np.allclose(res[1], e_o, rtol=1e-13, atol=0)
    assert np.allclose(res[0], p_o, rtol=1e-6, equal_nan=True)
