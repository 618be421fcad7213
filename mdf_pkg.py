"""Import helper: registers `metagenomic-deepfri_b200/` as module `metagenomic_deepfri_b200`."""
import importlib.util
import pathlib
import sys

NAME = "metagenomic_deepfri_b200"
ROOT = pathlib.Path(__file__).resolve().parent
PKG_DIR = ROOT / "metagenomic-deepfri_b200"


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(
        NAME, PKG_DIR / "__init__.py", submodule_search_locations=[str(PKG_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod
