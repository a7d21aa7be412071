"""binning.detect_edge_rule against stand-ins for the two powerbox generations (oracle with a fixed rule)."""
import sys
import types

import numpy as np
import pytest

import py21cmmc_b200  # noqa: F401
from py21cmmc_b200 import binning, likelihood
from oracle import ps_oracle as O


def _fake_powerbox(monkeypatch, rule):
    def get_power(box, boxlength, bins=None, bin_ave=True, get_variance=False, log_bins=False, k_weights=1, **kw):
        assert bin_ave is False and get_variance is False
        return O.get_power(box, boxlength, bins=bins, log_bins=log_bins, k_weights=k_weights, edge_rule=rule)

    pkg, tools = types.ModuleType("powerbox"), types.ModuleType("powerbox.tools")
    tools.get_power = get_power
    pkg.tools = tools
    monkeypatch.setitem(sys.modules, "powerbox", pkg)
    monkeypatch.setitem(sys.modules, "powerbox.tools", tools)


@pytest.mark.parametrize("rule", binning.EDGE_RULES)
def test_detects_generation(monkeypatch, rule):
    _fake_powerbox(monkeypatch, rule)
    assert binning.detect_edge_rule() == rule
    old = dict(likelihood._DEFAULTS)
    try:
        likelihood.set_defaults(edge_rule="auto")
        assert likelihood._DEFAULTS["edge_rule"] == rule
    finally:
        likelihood._DEFAULTS.update(old)


def test_unknown_generation_is_loud(monkeypatch):
    _fake_powerbox(monkeypatch, "legacy_1.001")
    tools = sys.modules["powerbox.tools"]
    inner = tools.get_power
    tools.get_power = lambda *a, **k: (inner(*a, **k)[0], inner(*a, **k)[1] * 1.01)
    with pytest.raises(RuntimeError):
        binning.detect_edge_rule()


def test_absent_powerbox_means_none(monkeypatch):
    monkeypatch.setitem(sys.modules, "powerbox", None)
    monkeypatch.setitem(sys.modules, "powerbox.tools", None)
    assert binning.detect_edge_rule() is None
