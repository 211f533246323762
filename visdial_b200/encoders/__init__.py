"""Encoder plugins, loaded by name like `dofile('encoders/<name>.lua')` (model.lua:19-20)."""
import importlib


def load(name: str):
    try:
        return importlib.import_module(__name__ + "." + name.replace("-", "_"))
    except ModuleNotFoundError:
        raise ValueError("unknown encoder '%s' (no encoders/%s)" % (name, name))
