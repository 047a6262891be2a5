"""Import helper: the package directory is named `mp3-enc-bsd_b200/` (not a valid Python identifier),
so it is loaded by path and registered as `mp3enc_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "mp3-enc-bsd_b200")


def load():
    if "mp3enc_b200" in sys.modules:
        return sys.modules["mp3enc_b200"]
    spec = importlib.util.spec_from_file_location("mp3enc_b200", os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["mp3enc_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
