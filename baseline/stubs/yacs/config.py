import copy

import yaml


class CfgNode(dict):
    """Attribute dictionary with yacs' merge helpers (immutability after freeze() is recorded, not enforced)."""

    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        super().__init__()
        object.__setattr__(self, "_frozen", False)
        object.__setattr__(self, "_new_allowed", new_allowed)
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        del self[name]

    def __deepcopy__(self, memo):
        out = CfgNode(new_allowed=self._new_allowed)
        for k, v in self.items():
            out[k] = copy.deepcopy(v, memo)
        return out

    def clone(self):
        return copy.deepcopy(self)

    def is_frozen(self):
        return self._frozen

    def _set_frozen(self, flag):
        object.__setattr__(self, "_frozen", flag)
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(flag)

    def freeze(self):
        self._set_frozen(True)

    def defrost(self):
        self._set_frozen(False)

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), CfgNode):
                self[k]._merge(v)
            else:
                self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def merge_from_other_cfg(self, other):
        self._merge(other)

    def merge_from_file(self, path):
        with open(path, "r") as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_list(self, items):
        assert len(items) % 2 == 0
        for key, value in zip(items[0::2], items[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                node = node[p]
            if isinstance(value, str):
                try:
                    value = yaml.safe_load(value)
                except yaml.YAMLError:
                    pass
            node[parts[-1]] = value

    def _plain(self):
        return {k: (v._plain() if isinstance(v, CfgNode) else v) for k, v in self.items()}

    def dump(self, **kwargs):
        return yaml.safe_dump(self._plain(), **kwargs)

    def __str__(self):
        return self.dump()
