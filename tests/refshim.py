"""Import the UNMODIFIED reference package from the git-ignored `baseline/_ref/` (put there by tools/vendor_reference.py) in an
image that lacks some of the packages it imports.  Nothing of the reference is edited; the stubs only satisfy imports of
things that are off the hot path:

  torchmetrics.detection.MeanAveragePrecision   detr/utils.py:3          (validation metrics)
  matplotlib.pyplot                             detr/visualize.py:4      (plots)
  accelerate, accelerate.utils                  detr/train.py:13-14      -> `accelerate_shim` (the ~16 members train.py touches)
  detr.model.get_model(weights="DEFAULT")       detr/model.py:432        -> weights=None (no network; BASELINE: random init)

`import_reference()` returns the imported `detr` package, or None when baseline/_ref is absent."""
import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference(with_train: bool = False):
    if not os.path.isfile(os.path.join(REF_DIR, "detr", "model.py")):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        import torchmetrics  # noqa: F401
    except Exception:
        tmd = _stub("torchmetrics.detection", MeanAveragePrecision=object)
        _stub("torchmetrics", detection=tmd)
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        plt = _stub("matplotlib.pyplot")
        _stub("matplotlib", pyplot=plt)
    if with_train:
        try:
            import accelerate  # noqa: F401
        except Exception:
            import accelerate_shim
            sys.modules["accelerate"] = accelerate_shim
            sys.modules["accelerate.utils"] = accelerate_shim.utils
    detr = importlib.import_module("detr")
    model = importlib.import_module("detr.model")
    importlib.import_module("detr.loss")
    importlib.import_module("detr.matcher")
    if not getattr(model, "_b200_offline", False):
        orig = model.get_model

        def get_model_offline(name, weights=None, **kw):   # detr/model.py:432 asks for pretrained weights: no network here
            return orig(name, weights=None, **kw)

        model.get_model = get_model_offline
        model._b200_offline = True
    if with_train:
        importlib.import_module("detr.train")
    return detr


def reload_reference_model():
    """A fresh, unpatched `detr.model` / `detr.loss` / `detr.matcher` (undoes detr_b200.model.patch)."""
    for name in ("detr.model", "detr.matcher", "detr.loss", "detr.train"):
        if name in sys.modules:
            importlib.reload(sys.modules[name])
    model = sys.modules["detr.model"]
    orig = model.get_model
    model.get_model = lambda name, weights=None, **kw: orig(name, weights=None, **kw)
    model._b200_offline = True
