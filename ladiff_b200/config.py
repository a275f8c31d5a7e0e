"""Config layer: the reference's plugin boundary.

The reference selects every hot-path component through ``instantiate_from_config`` (``src/ladiff/config.py:26-33``):
a node ``{target: "pkg.mod.Class", params: {...}}`` becomes ``Class(**params)``.  Its YAMLs are merged by OmegaConf
(``config.py:180-184``: ``base.yaml`` <- experiment cfg <- ``configs/<model.target>/*.yaml`` <- assets) and use
``${a.b.c}`` interpolation.  OmegaConf is not available here, so this module re-implements the three features the hot
path needs -- deep merge, attribute access and ``${...}`` interpolation -- on top of PyYAML, and adds ``retarget`` which
maps the reference's ``target:`` strings onto the B200 classes so the reference's own YAML files work unchanged.
"""
from __future__ import annotations

import copy
import importlib
import os
import re
from typing import Any, Iterable, Mapping

import yaml

# reference target -> B200-native drop-in
TARGET_MAP = {
    "ladiff.models.architectures.ladiff_denoiser.LADiffDenoiser": "ladiff_b200.denoiser.LADiffDenoiser",
    "ladiff.models.architectures.ladiff_vae.LADiffVae": "ladiff_b200.vae.LADiffVae",
    "ladiff.models.architectures.mld_clip.MldTextEncoder": "ladiff_b200.text_encoder.MldTextEncoder",
    "diffusers.DDIMScheduler": "ladiff_b200.scheduler.DDIMScheduler",
    "diffusers.DDPMScheduler": "ladiff_b200.scheduler.DDPMScheduler",
}


class Cfg(dict):
    """dict with attribute access and lazy ``${a.b}`` interpolation against the root node (OmegaConf subset)."""

    _INTERP = re.compile(r"\$\{([^}]+)\}")

    def __init__(self, data: Mapping = (), root: "Cfg" = None):
        super().__init__()
        object.__setattr__(self, "_root", root if root is not None else self)
        for k, v in dict(data).items():
            super().__setitem__(k, self._wrap(v))

    def _wrap(self, v):
        root = object.__getattribute__(self, "_root")
        if isinstance(v, Cfg):
            return Cfg(dict.copy(v), root)
        if isinstance(v, Mapping):
            return Cfg(v, root)
        return v

    def _resolve(self, v):
        if isinstance(v, str):
            root = object.__getattribute__(self, "_root")
            m = self._INTERP.fullmatch(v)
            if m:  # whole-value interpolation keeps the referenced type (lists, nodes, numbers)
                return root.select(m.group(1))
            return self._INTERP.sub(lambda mm: str(root.select(mm.group(1))), v)
        return v

    def select(self, dotted: str):
        node: Any = object.__getattribute__(self, "_root")
        for part in dotted.split("."):
            if not isinstance(node, Mapping) or part not in node:
                raise KeyError(f"interpolation key '{dotted}' not found")
            node = node[part]
        return node

    def __getitem__(self, k):
        return self._resolve(super().__getitem__(k))

    def get(self, k, default=None):
        return self[k] if k in self else default

    def __getattr__(self, k):
        if k not in self:
            raise AttributeError(k)
        return self[k]          # a missing ${...} target surfaces as KeyError, like OmegaConf's interpolation error

    def __setattr__(self, k, v):
        self[k] = v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]

    def to_dict(self) -> dict:
        out = {}
        for k in self.keys():
            v = self[k]
            out[k] = v.to_dict() if isinstance(v, Cfg) else copy.deepcopy(v)
        return out


def _raw(node):
    if isinstance(node, Mapping):
        return {k: _raw(dict.__getitem__(node, k)) for k in node.keys()}
    return node


def merge(*nodes: Mapping) -> Cfg:
    """Deep merge, later nodes win (OmegaConf.merge semantics for dict nodes; lists are replaced)."""
    def rec(a: dict, b: Mapping):
        for k, v in b.items():
            if isinstance(v, Mapping) and isinstance(a.get(k), Mapping):
                rec(a[k], v)
            else:
                a[k] = copy.deepcopy(v)
    out: dict = {}
    for n in nodes:
        rec(out, _raw(n))
    return Cfg(out)


def load_yaml(path: str) -> dict:
    with open(path, "r") as f:
        return yaml.safe_load(f) or {}


def get_module_config(cfg: Mapping, config_dir: str, path: str = "modules") -> Cfg:
    """``src/ladiff/config.py:7-13``: every ``configs/<path>/*.yaml`` is merged into ``cfg.model``."""
    d = os.path.join(config_dir, path)
    model = merge(cfg.get("model", {}) if isinstance(cfg, Mapping) else {},
                  *[load_yaml(os.path.join(d, f)) for f in sorted(os.listdir(d)) if f.endswith(".yaml")])
    return merge(cfg, {"model": model})


def load_config(cfg_path: str, config_dir: str = None, assets_path: str = None, base_path: str = None,
                overrides: Mapping = None, retarget_refs: bool = True) -> Cfg:
    """Merge order of ``parse_args`` (``src/ladiff/config.py:180-184``):
    base.yaml <- --cfg <- configs/<model.target>/*.yaml <- --cfg_assets (<- overrides)."""
    config_dir = config_dir or os.path.dirname(os.path.abspath(cfg_path))
    base_path = base_path or os.path.join(config_dir, "base.yaml")
    cfg = merge(load_yaml(base_path) if os.path.exists(base_path) else {}, load_yaml(cfg_path))
    target = cfg.get("model", {}).get("target", "modules") if "model" in cfg else "modules"
    cfg = get_module_config(cfg, config_dir, target)
    if assets_path and os.path.exists(assets_path):
        cfg = merge(cfg, load_yaml(assets_path))
    if overrides:
        cfg = merge(cfg, overrides)
    return retarget(cfg) if retarget_refs else cfg


def retarget(cfg: Mapping) -> Cfg:
    """Rewrites reference ``target:`` strings to the B200 drop-ins (only the hot-path components are mapped)."""
    def rec(n):
        if isinstance(n, dict):
            for k, v in list(n.items()):
                if k == "target" and isinstance(v, str) and v in TARGET_MAP:
                    n[k] = TARGET_MAP[v]
                else:
                    rec(v)
        elif isinstance(n, list):
            for v in n:
                rec(v)
    raw = _raw(cfg)
    rec(raw)
    return Cfg(raw)


def get_obj_from_str(string: str):
    module, cls = string.rsplit(".", 1)
    return getattr(importlib.import_module(module, package=None), cls)


def instantiate_from_config(config: Mapping, **extra):
    """``src/ladiff/config.py:26-33`` (same error behaviour: KeyError without ``target``)."""
    if "target" not in config:
        if config == "__is_first_stage__" or config == "__is_unconditional__":
            return None
        raise KeyError("Expected key `target` to instantiate.")
    params = config.get("params", dict())
    params = params.to_dict() if isinstance(params, Cfg) else dict(params)
    # keep attribute-style access for the ablation node (the reference passes an OmegaConf node)
    if isinstance(config.get("params", None), Cfg) and "ablation" in config["params"]:
        params["ablation"] = config["params"]["ablation"]
    params.update(extra)
    return get_obj_from_str(config["target"])(**params)
