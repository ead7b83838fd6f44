"""Minimal Hydra/OmegaConf stand-in for the reference's config tree (neither library is installed here).

Supports exactly what config/config.yaml and config/inference.yaml use: a ``defaults`` list of config groups
(``- model: default`` -> config/model/default.yaml under key ``model``), ``_self_``, ``override hydra/...``
entries (ignored), ``${a.b}`` interpolation, ``${now:%fmt}``, and ``a.b=value`` command-line overrides.
When hydra IS importable the reference entry points may use it instead; the resulting tree has the same keys.
"""
from __future__ import annotations

import datetime
import os
import re
from typing import Any, Dict, Iterable, List, Optional

import yaml


class Cfg(dict):
    """dict with attribute access (the subset of DictConfig the reference code relies on)"""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_cfg(obj: Any) -> Any:
    if isinstance(obj, dict):
        return Cfg({k: to_cfg(v) for k, v in obj.items()})
    if isinstance(obj, list):
        return [to_cfg(v) for v in obj]
    return obj


def to_container(obj: Any) -> Any:
    if isinstance(obj, dict):
        return {k: to_container(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [to_container(v) for v in obj]
    return obj


def _merge(dst: Dict, src: Dict) -> Dict:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def _lookup(root: Dict, dotted: str):
    cur: Any = root
    for part in dotted.split("."):
        cur = cur[part]
    return cur


_INTERP = re.compile(r"\$\{([^${}]+)\}")


def _resolve(node: Any, root: Dict, now: datetime.datetime) -> Any:
    if isinstance(node, dict):
        for k in list(node):
            node[k] = _resolve(node[k], root, now)
        return node
    if isinstance(node, list):
        return [_resolve(v, root, now) for v in node]
    if isinstance(node, str) and "${" in node:
        def sub(m):
            expr = m.group(1)
            if expr.startswith("now:"):
                return now.strftime(expr[4:])
            try:
                return str(_resolve(_lookup(root, expr), root, now))
            except (KeyError, TypeError):
                return m.group(0)  # e.g. ${hydra.job.override_dirname}: left for hydra itself
        whole = _INTERP.fullmatch(node)
        if whole and not whole.group(1).startswith("now:"):
            try:
                return _resolve(_lookup(root, whole.group(1)), root, now)
            except (KeyError, TypeError):
                return node
        return _INTERP.sub(sub, node)
    return node


def _parse_value(text: str) -> Any:
    return yaml.safe_load(text)


def compose(config_dir: str, config_name: str = "config", overrides: Optional[Iterable[str]] = None) -> Cfg:
    with open(os.path.join(config_dir, config_name + ".yaml")) as f:
        primary = yaml.safe_load(f) or {}
    defaults: List[Any] = primary.pop("defaults", ["_self_"])
    if "_self_" not in defaults:
        defaults = list(defaults) + ["_self_"]  # hydra >= 1.1: the primary config is merged last by default
    out: Dict[str, Any] = {}
    for entry in defaults:
        if entry == "_self_":
            _merge(out, primary)
        elif isinstance(entry, dict):
            (group, option), = entry.items()
            if str(group).startswith("override ") or str(group).startswith("hydra/"):
                continue
            with open(os.path.join(config_dir, group, f"{option}.yaml")) as f:
                _merge(out.setdefault(group, {}), yaml.safe_load(f) or {})
    for ov in overrides or []:
        if "=" not in ov:
            raise ValueError(f"override '{ov}' is not of the form key=value")
        key, val = ov.split("=", 1)
        key = key.lstrip("+")
        cur = out
        parts = key.split(".")
        for p in parts[:-1]:
            cur = cur.setdefault(p, {})
        cur[parts[-1]] = _parse_value(val)
    _resolve(out, out, datetime.datetime.now())
    return to_cfg(out)
