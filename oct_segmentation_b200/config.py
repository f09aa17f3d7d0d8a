"""Minimal Hydra-compatible config composition for the predict entry point.

The reference uses ``@hydra.main(config_path=PROJECT_DIR/configs, config_name='predict')``
(/root/reference/src/predict.py:104-108) with ``defaults: [main, _self_]``
(/root/reference/configs/predict.yaml:1-3).  Hydra/OmegaConf are not available offline, so this
module composes the same files with PyYAML: defaults list (in order, ``_self_`` = the file
itself), ``key=value`` / ``key=[a,b]`` / ``a.b=value`` command-line overrides, attribute access.
"""
from __future__ import annotations

import os
from typing import Any, Dict, Iterable, List

import yaml


class Config(dict):
    """dict with attribute access (cfg.device, cfg.classes ...)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return v

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(x: Any) -> Any:
    if isinstance(x, dict):
        return Config({k: _wrap(v) for k, v in x.items()})
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def _merge(dst: Dict, src: Dict) -> Dict:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def _load_yaml(path: str) -> Dict:
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    with open(path) as f:
        return yaml.safe_load(f) or {}


def compose(config_dir: str, config_name: str, overrides: Iterable[str] = ()) -> Config:
    body = _load_yaml(os.path.join(config_dir, config_name + '.yaml'))
    defaults: List = body.pop('defaults', ['_self_'])
    if '_self_' not in defaults:
        defaults = list(defaults) + ['_self_']
    out: Dict = {}
    for d in defaults:
        if d == '_self_':
            _merge(out, body)
        else:
            name = d if isinstance(d, str) else next(iter(d.values()))
            sub = compose(config_dir, name)
            _merge(out, dict(sub))
    for ov in overrides:
        if '=' not in ov:
            raise ValueError(f'override `{ov}` is not of the form key=value')
        key, val = ov.split('=', 1)
        key = key.lstrip('+')
        node = out
        parts = key.split('.')
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = yaml.safe_load(val)
    return _wrap(out)


def to_yaml(cfg: Dict) -> str:
    def plain(x):
        if isinstance(x, dict):
            return {k: plain(v) for k, v in x.items()}
        if isinstance(x, list):
            return [plain(v) for v in x]
        return x
    return yaml.safe_dump(plain({k: v for k, v in cfg.items() if k != 'hydra'}), sort_keys=False)
