"""The host layer keeps the reference's public Python surface: signatures, registry keys and CLI flags, checked
against tests/golden/api_surface.json (extracted from the reference with `ast` by tests/golden/make_surface_golden.py)."""
import inspect
import json
import os

import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "api_surface.json")))


def _sig(fn, skip_self=False):
    out = []
    for name, p in inspect.signature(fn).parameters.items():
        if skip_self and name == "self":
            continue
        if p.kind is inspect.Parameter.VAR_KEYWORD:
            out.append({"name": "**" + name, "has_default": False, "default": None})
        else:
            has = p.default is not inspect.Parameter.empty
            d = p.default if has else None
            out.append({"name": name, "has_default": has, "default": list(d) if isinstance(d, tuple) else d})
    return out


def _norm(params):
    return [{"name": p["name"], "has_default": p["has_default"],
             "default": list(p["default"]) if isinstance(p["default"], (list, tuple)) else p["default"]} for p in params]


def test_api_dca_signature_is_the_reference_signature():
    from dca_b200 import api
    assert _sig(api.dca) == _norm(GOLD["api.dca"])


def test_train_signature_is_the_reference_signature():
    from dca_b200 import train
    assert _sig(train.train) == _norm(GOLD["train.train"])


def test_autoencoder_constructor_keeps_the_reference_parameters():
    """network.py:44-59; extra keyword parameters of this implementation may only follow the reference's."""
    from dca_b200.network import AE_types
    ours = _sig(AE_types["zinb-conddisp"].__init__, skip_self=True)
    ref = [p for p in _norm(GOLD["network.Autoencoder.__init__"]) if p["name"] != "self"]
    assert ours[:len(ref)] == ref
    assert all(p["has_default"] or p["name"].startswith("**") for p in ours[len(ref):])


def test_registry_has_every_reference_key():
    from dca_b200.network import AE_types
    assert list(AE_types.keys()) == GOLD["network.AE_types"]


def test_cli_accepts_every_reference_flag_with_the_same_default():
    from dca_b200.__main__ import build_parser
    parser = build_parser()
    ours = {}
    for act in parser._actions:
        for s in act.option_strings or [act.dest]:
            ours[s] = act
    for flag in GOLD["cli"]:
        for name in flag["names"]:
            assert name in ours, "missing CLI flag %s" % name
        act = ours[flag["names"][0]]
        if "default" in flag and not str(flag["default"]).startswith("<expr:"):
            assert act.default == flag["default"], (flag["names"], act.default, flag["default"])
        if "dest" in flag:
            assert act.dest == flag["dest"], flag["names"]
