"""YAML configuration tree with the reference's schema (`diy_gym/config.py:13-61`).

A nested mapping is a *model* iff it holds the key `model` and an *addon* iff it holds the key
`addon`; the environment name defaults to the file stem.  Behaviour kept verbatim: `get` wraps nested
mappings in a `Configuration`, a missing key without default raises `KeyError` with the reference's
wording, unknown keys are ignored.
"""
import os

import yaml

_MISSING = object()


class Configuration:
    def __init__(self, name, node):
        self.name = name
        self.node = node

    @classmethod
    def from_file(cls, path):
        with open(path, 'r') as stream:
            top = yaml.load(stream, Loader=yaml.FullLoader)
        if top is None:
            top = {}
        stem = os.path.splitext(os.path.basename(path))[0]
        return cls(top['name'] if 'name' in top else stem, top)

    @classmethod
    def from_dict(cls, name, node):
        return cls(name, node)

    def get(self, key, default=_MISSING):
        if key in self.node:
            val = self.node[key]
            return Configuration(key, val) if isinstance(val, dict) else val
        if default is not _MISSING:
            return default
        raise KeyError("Couldn't find config and no default provided for config with key: " + key)

    def set(self, key, val):
        self.node[key] = val

    def __contains__(self, key):
        return key in self.node

    def find_all(self, key):
        """Yield every directly nested config that contains `key` (YAML order)."""
        for k, v in self.node.items():
            if isinstance(v, dict) and key in v:
                yield Configuration(k, v)

    def find(self, key):
        return next(iter(self.find_all(key)))
