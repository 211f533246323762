"""Decoder plugins, loaded by name like `dofile('decoders/<name>.lua')` (model.lua:22-23)."""
import importlib


def load(name: str):
    try:
        return importlib.import_module(__name__ + "." + name)
    except ModuleNotFoundError:
        raise ValueError("unknown decoder '%s' (no decoders/%s)" % (name, name))
