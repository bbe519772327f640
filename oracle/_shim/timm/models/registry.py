"""Stand-in for timm.models.registry.register_model (networks/convnet.py:5 decorates its factories with it; the
RDST path never calls them)."""


def register_model(fn):
    return fn
