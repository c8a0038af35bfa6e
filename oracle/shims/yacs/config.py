"""Minimal stand-in for yacs.config.CfgNode (yacs is not installed in this image).

Only what the reference's config/defaults.py and tools/test_net.py use: attribute access,
nested nodes, merge_from_file / merge_from_list, freeze / defrost, clone.
"""
import ast
import copy

import yaml


class CfgNode(dict):
    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        super().__init__()
        self.__dict__["_frozen"] = False
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.__dict__.get("_frozen", False):
            raise AttributeError("attempted to modify frozen CfgNode: %s" % name)
        self[name] = value

    def freeze(self):
        self._set_frozen(True)

    def defrost(self):
        self._set_frozen(False)

    def _set_frozen(self, flag):
        self.__dict__["_frozen"] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(flag)

    def is_frozen(self):
        return self.__dict__["_frozen"]

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        new = CfgNode()
        for k, v in self.items():
            dict.__setitem__(new, k, copy.deepcopy(v, memo))
        new.__dict__["_frozen"] = self.__dict__["_frozen"]
        return new

    @staticmethod
    def _coerce(new, old):
        if isinstance(old, tuple) and isinstance(new, list):
            return tuple(new)
        if isinstance(old, list) and isinstance(new, tuple):
            return list(new)
        if isinstance(old, float) and isinstance(new, int):
            return float(new)
        return new

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict):
                if k not in self:
                    self[k] = CfgNode()
                self[k]._merge(v)
            else:
                if isinstance(v, str):
                    try:
                        v = ast.literal_eval(v)
                    except (ValueError, SyntaxError):
                        pass
                self[k] = self._coerce(v, self.get(k)) if k in self else v

    def merge_from_file(self, path):
        with open(path) as fh:
            data = yaml.safe_load(fh) or {}
        self._merge(data)

    def merge_from_other_cfg(self, other):
        self._merge(other)

    def merge_from_list(self, lst):
        assert len(lst) % 2 == 0
        for full_key, v in zip(lst[0::2], lst[1::2]):
            node = self
            parts = full_key.split(".")
            for p in parts[:-1]:
                node = node[p]
            if isinstance(v, str):
                try:
                    v = ast.literal_eval(v)
                except (ValueError, SyntaxError):
                    pass
            node[parts[-1]] = self._coerce(v, node.get(parts[-1]))
