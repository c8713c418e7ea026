"""Import the UNMODIFIED reference (/root/reference) in the authoring container.  TEST INFRASTRUCTURE ONLY.

/root/reference does not exist on the GPU box, so nothing that runs there may import this module; it is used by
`oracle/make_golden.py` (golden fixture generation), `tests/test_oracle_vs_reference.py` (skipped when the reference
is absent) and nothing else.

The reference's `models/__init__.py` imports timm (absent), so the package __init__ is bypassed by registering an
empty `models` package whose __path__ points at the reference directory.
"""
import os
import sys
import types

# Where the reference lives: the mounted tree in the authoring container; on the GPU box (no /root/reference) the
# git-ignored copy `oracle/_ref/` that `oracle/stage_reference.py` (run by __graft_entry__.build() here) stages so that
# bench.py's reference arm can time the UNMODIFIED reference modules on the box's host cores.
_HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(_HERE, "_ref")


def _default_root():
    if os.path.isfile(os.path.join("/root/reference", "models", "timesformer.py")):
        return "/root/reference"
    return _STAGED


REF_ROOT = os.environ.get("DVST_REFERENCE_ROOT") or _default_root()


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "models", "timesformer.py"))


class _Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def make_cfg(crop=224, num_frames=8, num_classes=400, attention_type="divided_space_time"):
    """Attribute bag with the five fields get_vit_base_patch16_224 reads (models/timesformer.py:594-602)."""
    return _Bag(DATA=_Bag(TRAIN_CROP_SIZE=crop, NUM_FRAMES=num_frames), MODEL=_Bag(NUM_CLASSES=num_classes),
                TIMESFORMER=_Bag(ATTENTION_TYPE=attention_type, PRETRAINED_MODEL=""))


_loaded = {}


def load():
    """Returns a dict of reference symbols."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if "models" not in sys.modules or not hasattr(sys.modules["models"], "__path__"):
        pkg = types.ModuleType("models")
        pkg.__path__ = [os.path.join(REF_ROOT, "models")]
        sys.modules["models"] = pkg
    import importlib
    tsf = importlib.import_module("models.timesformer")
    sys.modules["models"].get_vit_base_patch16_224 = tsf.get_vit_base_patch16_224
    vt = importlib.import_module("vision_transformer")

    # dino_similarity.py imports datasets_custom / utils / fvcore at module level: stub what is absent.
    if "fvcore" not in sys.modules:
        for name in ("fvcore", "fvcore.common", "fvcore.common.file_io"):
            sys.modules[name] = types.ModuleType(name)
        sys.modules["fvcore.common.file_io"].PathManager = object
    if "datasets_custom" not in sys.modules:
        dc = types.ModuleType("datasets_custom")
        dc.__path__ = [os.path.join(REF_ROOT, "datasets_custom")]
        sys.modules["datasets_custom"] = dc
        # only torch/torchvision are needed by dino_loss_loader once transform / data_utils import
        try:
            dll = importlib.import_module("datasets_custom.dino_loss_loader")
            dc.DinoLossLoader = dll.DinoLossLoader
        except Exception:  # pragma: no cover - optional
            dll = None
            dc.DinoLossLoader = object
    else:
        dll = sys.modules.get("datasets_custom.dino_loss_loader")
    # `utils` : path-only package; utils.parser (needs fvcore/yacs) is replaced by an identity load_config.
    up = sys.modules.get("utils")
    if up is None or not hasattr(up, "__path__") or os.path.join(REF_ROOT, "utils") not in list(up.__path__):
        up = types.ModuleType("utils")
        up.__path__ = [os.path.join(REF_ROOT, "utils")]
        sys.modules["utils"] = up
    if "utils.utils" not in sys.modules:
        up.utils = types.ModuleType("utils.utils")
        sys.modules["utils.utils"] = up.utils
    parser = types.ModuleType("utils.parser")
    parser.load_config = lambda args: args
    sys.modules["utils.parser"] = parser
    up.parser = parser
    ds = importlib.import_module("dino_similarity")
    _loaded.update(timesformer=tsf, vision_transformer=vt, dino_similarity=ds, dino_loss_loader=dll,
                   get_vit_base_patch16_224=tsf.get_vit_base_patch16_224, VisionTransformer=tsf.VisionTransformer,
                   DINOHead=vt.DINOHead, DINOLoss=ds.DINOLoss, export_loss=ds.export_loss)
    return _loaded


def load_knn_classifier():
    """The reference's `knn_classifier` (eval_knn.py:138-178) compiled from its own source text.  eval_knn.py cannot be
    imported as a module (kornia / fvcore dependent dataset imports at the top), so the function definition is cut out
    with `ast` and executed unmodified in a namespace that holds only `torch`."""
    import ast
    import torch
    path = os.path.join(REF_ROOT, "eval_knn.py")
    src = open(path).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "knn_classifier")
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"torch": torch}
    exec(compile(mod, path, "exec"), ns)
    return ns["knn_classifier"]
