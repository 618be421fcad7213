import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)
import mdf_pkg  # noqa: E402

mdf_pkg.load()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cmap_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "cmap_golden.npz"))


@pytest.fixture(scope="session")
def gcn_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "gcn_golden.npz"))


@pytest.fixture(scope="session")
def model_dir(tmp_path_factory):
    """Seeded random-init ONNX heads of tests/golden/spec.py, written once per session."""
    import spec
    from metagenomic_deepfri_b200 import synth
    d = tmp_path_factory.mktemp("models")
    paths = {}
    for tag, (kw, seed, n, lo, hi) in spec.GCN_CASES.items():
        p = str(d / f"{tag}.onnx")
        synth.write_gcn_model(p, synth.GCNConfig(**kw), seed=seed)
        paths[tag] = p
    return paths


@pytest.fixture(scope="session")
def cnn_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "cnn_golden.npz"))


@pytest.fixture(scope="session")
def cnn_model_dir(tmp_path_factory):
    """Seeded random-init DeepCNN heads of tests/golden/spec.py (sequence-only branch)."""
    import spec
    from metagenomic_deepfri_b200 import synth
    d = tmp_path_factory.mktemp("cnn_models")
    paths = {}
    for tag, (kw, seed, n, lo, hi) in spec.CNN_CASES.items():
        p = str(d / f"{tag}.onnx")
        synth.write_cnn_model(p, synth.CNNConfig(**kw), seed=seed)
        paths[tag] = p
    return paths


def golden_workload(tag):
    import spec
    from metagenomic_deepfri_b200 import synth
    kw, seed, n, lo, hi = spec.GCN_CASES[tag]
    return synth.make_workload(n, lo, hi, seed=spec.workload_seed(seed), threshold=spec.THRESHOLD)


class Aln:
    """Minimal stand-in for mDeepFRI.alignment.AlignmentResult (alignment.py:65-150)."""

    def __init__(self, gq, gt, coords, i=0):
        self.query_name = f"query_{i}"
        self.target_name = f"target_{i}.pdb"
        self.gapped_sequence = gq
        self.gapped_target = gt
        self.query_sequence = gq.replace("-", "")
        self.coords = coords


class Recognised:
    """What the library's loader (`mdf_onnx_inspect` / `mdf_onnx_tensor`, host only) makes of an `.onnx` file."""

    def __init__(self, path, temporary=False):
        from metagenomic_deepfri_b200 import _lib
        self.path = path
        self.temporary = temporary
        try:
            self.info = _lib.inspect_onnx(path)
        except Exception:
            self.__del__()
            raise

    def __del__(self):
        if self.__dict__.get("temporary") and os.path.exists(self.path):
            os.unlink(self.path)

    def __getattr__(self, k):
        try:
            return self.__dict__["info"][k]
        except KeyError:
            raise AttributeError(k)

    def tensor(self, role, shape=None):
        from metagenomic_deepfri_b200 import _lib
        a = _lib.onnx_tensor(self.path, role)
        return a.reshape(shape) if shape is not None else a


def recognise(model):
    """`model`: a path, or an onnx_lite.Model (written to a temporary file first)."""
    import tempfile
    from metagenomic_deepfri_b200 import onnx_lite
    if isinstance(model, str):
        return Recognised(model)
    with tempfile.NamedTemporaryFile(suffix=".onnx", delete=False) as fh:
        fh.write(onnx_lite.dumps(model))
    return Recognised(fh.name, temporary=True)
