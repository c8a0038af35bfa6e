"""name -> constructor registries (same role as the reference's utils/registry.py)"""


class Registry(dict):
    def register(self, name, module=None):
        if module is not None:
            self[name] = module
            return module

        def deco(fn):
            self[name] = fn
            return fn
        return deco
