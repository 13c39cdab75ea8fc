"""`src.datasets` for the drop-in `src` package.

Datasets, augmentation and annotation loading are outside the hot path (SURVEY.md section 8 / DESIGN.md section 7): they are
host-side Python that this build does not replace.  So that the reference's own entry points keep working unmodified --
`from src import datasets, models` and `datasets.__dict__[cfg['DATASET']['name']]` (scripts/train_and_evaluate.py:18,56;
src/runner/trainer.py:10,47-48) -- this package RE-EXPORTS the reference's dataset classes from a reference checkout when
one is present:

    HG_REFERENCE_SRC=/path/to/hourglass-pose-estimation/src      (or /root/reference/src when that exists)

Its modules (`common.py`, `mpii.py`, `mscoco.py`) are imported from there under this package's name; what they import
from `src.utils` resolves to this build's API-compatible modules.  `mscoco` needs `pycocotools`; a dataset whose
dependencies are missing is simply not exported.  Without a checkout the package is empty and `datasets.__dict__[name]`
raises the same KeyError an unknown dataset name gives in the reference.

The on-device replacement of `JointsDataset.generate_target` (common.py:197-248) lives in hgb200.ops
(`joint_centers` + `gaussian_target`, or `src.loss.JointsMSELossOnTheFly`).
"""
import importlib
import os

__all__ = []
REFERENCE_DATASETS = None          # directory the classes were taken from, or None


def _find():
    cands = [os.environ.get("HG_REFERENCE_SRC"), "/root/reference/src"]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "datasets", "common.py")):
            return os.path.join(c, "datasets")
    return None


_dir = _find()
if _dir is not None:
    __path__.append(_dir)
    REFERENCE_DATASETS = _dir
    for _name in ("mpii", "mscoco"):
        try:
            _mod = importlib.import_module(f"{__name__}.{_name}")
            globals()[_name] = getattr(_mod, _name)
            __all__.append(_name)
        except Exception as _e:          # e.g. pycocotools missing for mscoco: that dataset is not exported
            globals()[f"_{_name}_import_error"] = _e
    __all__ = tuple(__all__)
