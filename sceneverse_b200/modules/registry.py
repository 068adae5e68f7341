"""Registration under the reference's registry names (modules/build.py:6-22).

If the reference package is importable (a deployment where SceneVerse is on sys.path), the classes are
registered INTO the reference's own `VISION_REGISTRY` / `GROUNDING_REGISTRY` / `HEADS_REGISTRY`
(`install_into_reference()`), replacing the stock entries so that `model/build.py` and `trainer/`
consume them unchanged.  Standalone, the same `build_module(kind, cfg)` entry point is provided here.
"""


class Registry:
    def __init__(self, name):
        self._name, self._obj_map = name, {}

    def register(self, obj=None):
        def deco(o):
            self._obj_map[o.__name__] = o
            return o
        return deco if obj is None else deco(obj)

    def get(self, name):
        if name not in self._obj_map:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self._obj_map[name]

    def __contains__(self, name):
        return name in self._obj_map


VISION_REGISTRY = Registry("vision")
LANGUAGE_REGISTRY = Registry("language")
GROUNDING_REGISTRY = Registry("grounding")
HEADS_REGISTRY = Registry("heads")
LOSS_REGISTRY = Registry("loss")


def _args(cfg):
    a = getattr(cfg, "args", None)
    if a is None and isinstance(cfg, dict):
        a = cfg.get("args")
    return dict(a) if a else {}


def _name(cfg):
    return cfg["name"] if isinstance(cfg, dict) else cfg.name


def build_module(module_type, cfg):
    """modules/build.py:12-22."""
    reg = {"vision": VISION_REGISTRY, "language": LANGUAGE_REGISTRY, "grounding": GROUNDING_REGISTRY,
           "heads": HEADS_REGISTRY}.get(module_type)
    if reg is None:
        raise NotImplementedError(f"module type {module_type} not implemented")
    return reg.get(_name(cfg))(cfg, **_args(cfg))


def install_into_reference():
    """Overwrite the reference registries' entries with the B200 classes (call after `import modules`)."""
    import modules.build as ref  # the reference package
    from . import grounding, heads, vision  # noqa: F401  (populate the local registries)
    for mine, theirs in [(VISION_REGISTRY, ref.VISION_REGISTRY), (GROUNDING_REGISTRY, ref.GROUNDING_REGISTRY),
                         (HEADS_REGISTRY, ref.HEADS_REGISTRY)]:
        for name, cls in mine._obj_map.items():
            theirs._obj_map[name] = cls  # fvcore's Registry keeps its table in `_obj_map` too
    return ref
