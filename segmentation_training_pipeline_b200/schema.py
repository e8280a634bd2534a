"""The reference's own YAML schema as the authority on keys (reference segmentation.py:119-129: createNet1 forwards every
NON-custom schema property whose alias is a parameter of the architecture's constructor).  schemas/segmentation.raml is the
reference's file, kept verbatim as DATA (it is plain YAML); this module reads it to decide, per architecture, which keys are
MODEL keyword arguments, what they are called on the constructor side, and what their defaults are -- so that a key this
engine does not build is rejected BY NAME instead of being dropped."""
from __future__ import annotations

import os
from functools import lru_cache
from typing import Dict, Optional, Tuple

import yaml

_RAML = os.path.join(os.path.dirname(os.path.abspath(__file__)), "schemas", "segmentation.raml")


@lru_cache(maxsize=1)
def _types() -> dict:
    with open(_RAML) as f:
        return yaml.safe_load(f)["types"]


def _props(type_name: str) -> Dict[str, dict]:
    out = {}
    for k, v in (_types().get(type_name, {}).get("properties") or {}).items():
        out[k.rstrip("?")] = v if isinstance(v, dict) else {"type": v}
    return out


def architectures():
    return list(_props("PipelineConfig")["architecture"]["enum"])


def model_keys(architecture: str) -> Dict[str, Tuple[str, object]]:
    """{yaml key: (constructor keyword, schema default)} of the keys createNet1 would forward to the architecture: the
    non-custom properties of PipelineConfig and of the architecture's subtype."""
    out: Dict[str, Tuple[str, object]] = {}
    for tname in ("PipelineConfig", architecture):
        for k, v in _props(tname).items():
            if v.get("(meta.custom)"):
                continue
            out[k] = (v.get("(meta.alias)", k), v.get("default"))
    return out


def config_keys() -> set:
    """every key the schema knows for a PipelineConfig of any architecture (plus StageConfig's for `stages:` entries)"""
    ks = set(_props("PipelineConfig"))
    for a in architectures():
        ks |= set(_props(a))
    return ks


def stage_keys() -> set:
    return set(_props("StageConfig")) | set(_props("HasLoss"))


def resolve(architecture: str, given: Dict[str, object]) -> Dict[str, object]:
    """Model keyword arguments of `architecture` found in `given` (under the YAML key or under the constructor alias), keyed
    by YAML key, schema defaults filled in for absent ones."""
    mk = model_keys(architecture)
    out = {}
    for k, (alias, default) in mk.items():
        if k in given:
            out[k] = given[k]
        elif alias in given:
            out[k] = given[alias]
        else:
            out[k] = default
    return out
