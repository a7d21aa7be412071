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
        likelihood._AUTO_RULE.clear()
        likelihood.set_defaults(edge_rule="auto")
        assert likelihood.resolve_edge_rule() == rule            # probed once per process ...
        assert likelihood.resolve_edge_rule("auto") == rule      # ... also for instances built with the default
        assert likelihood._AUTO_RULE == [rule]
        likelihood.set_defaults(edge_rule="inscribed")           # an explicit process default wins over probing
        assert likelihood.resolve_edge_rule("auto") == "inscribed"
        assert likelihood.resolve_edge_rule("legacy_1.001") == "legacy_1.001"
    finally:
        likelihood._DEFAULTS.update(old)
        likelihood._AUTO_RULE.clear()


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


def test_auto_without_powerbox_falls_back_to_legacy_with_a_warning(monkeypatch, caplog):
    monkeypatch.setitem(sys.modules, "powerbox", None)
    monkeypatch.setitem(sys.modules, "powerbox.tools", None)
    old = dict(likelihood._DEFAULTS)
    try:
        likelihood._AUTO_RULE.clear()
        likelihood.set_defaults(edge_rule="auto")
        with caplog.at_level("WARNING", logger="21cmFAST"):
            assert likelihood.resolve_edge_rule() == "legacy_1.001"
        assert any("powerbox is not importable" in r.message for r in caplog.records)
        lk = likelihood.Likelihood1DPowerLightcone(nchunks=2, use_data=False)   # the constructor default is "auto"
        assert lk.edge_rule == "auto"
        with pytest.raises(ValueError):
            likelihood.Likelihood1DPowerLightcone(nchunks=2, edge_rule="newest")
    finally:
        likelihood._DEFAULTS.update(old)
        likelihood._AUTO_RULE.clear()
