"""Import shims for running the UNMODIFIED reference (luoao-kddi/SCP) from /root/reference.

TEST INFRASTRUCTURE ONLY.  This module is used by ``oracle/make_golden.py`` in the build
container (where /root/reference exists) to produce the fixtures under ``tests/golden``.
It is never imported by the product (``scp_b200``), by ``-m gpu`` tests, by ``smoke()`` or
by ``bench.py`` -- /root/reference does not exist on the GPU box.

The reference cannot be imported as-is here because a few of its imports are not installed
(SURVEY.md section 8c):
  * ``pytorch_lightning``  -> stub whose ``LightningModule`` is ``torch.nn.Module``
  * two names from ``transformers`` that swin_transformer.py imports but never calls
  * ``h5py`` / ``plyfile`` / ``hydra`` / ``tqdm`` (present) / ``open3d``
No reference source is copied: the modules are executed from where they lie.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("SCP_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "encode.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_shims():
    import torch

    if "pytorch_lightning" not in sys.modules:
        class LightningModule(torch.nn.Module):
            def log(self, *a, **k):
                pass

            @classmethod
            def load_from_checkpoint(cls, *a, **k):
                raise RuntimeError("no checkpoints offline")

        _stub("pytorch_lightning", LightningModule=LightningModule)
    import transformers.pytorch_utils as tpu
    if not hasattr(tpu, "find_pruneable_heads_and_indices"):
        tpu.find_pruneable_heads_and_indices = lambda *a, **k: None
    try:
        import transformers.utils.backbone_utils as bu
    except Exception:  # pragma: no cover
        bu = _stub("transformers.utils.backbone_utils")
    if not hasattr(bu, "get_aligned_output_features_output_indices"):
        bu.get_aligned_output_features_output_indices = lambda *a, **k: (None, None)
    for missing in ("h5py", "open3d"):
        if missing not in sys.modules:
            try:
                importlib.import_module(missing)
            except Exception:
                _stub(missing)
    if "plyfile" not in sys.modules:
        try:
            importlib.import_module("plyfile")
        except Exception:
            _stub("plyfile", PlyData=object, PlyElement=object)
    if "hydra" not in sys.modules:
        try:
            importlib.import_module("hydra")
        except Exception:
            _stub("hydra", initialize=lambda *a, **k: None, compose=lambda *a, **k: None)


def import_reference():
    """Returns a namespace with the reference modules needed for golden generation."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    install_shims()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    cwd = os.getcwd()
    ns = types.SimpleNamespace()
    try:
        ns.Octree = importlib.import_module("data_preproc.Octree")
        ns.Octreewarpper = importlib.import_module("data_preproc.OctreeCPP.Octreewarpper")
        ns.data_preprocess = importlib.import_module("data_preproc.data_preprocess")
        ns.ehem = importlib.import_module("models.ehem")
        ns.oct_attention = importlib.import_module("models.oct_attention")
        ns.ds_ehem = importlib.import_module("dataloaders.encode_dataset_ehem")
        ns.ds_ehem_mul = importlib.import_module("dataloaders.encode_dataset_ehem_mullevel")
        ns.ds_oct = importlib.import_module("dataloaders.encode_dataset")
        ns.ds_oct_mul = importlib.import_module("dataloaders.encode_dataset_mullevel")
    finally:
        os.chdir(cwd)
    return ns


def make_cfg(model="ehem", train_type="kitti"):
    """cfg namespace with the values of the reference's configs/model/*.yaml."""
    NS = types.SimpleNamespace
    if model == "ehem":
        m = NS(class_name="EHEM", context_size=8192, token_num=255, layer_num=3, head_num=4,
               abs_pos_embed_dim=0, occ_embed_dim=54, level_embed_dim=6, octant_embed_dim=4,
               hidden_dimension=300, pos_max_len=5000, level_k=4, pos_embed=True, max_level=19)
    else:
        m = NS(class_name="OctAttention", max_octree_level=12, context_size=1024, token_num=255,
               layer_num=3, head_num=4, abs_pos_embed_dim=12, occ_embed_dim=128, level_embed_dim=6,
               octant_embed_dim=4, hidden_dimension=300, pos_max_len=5000, level_k=4, pos_embed=True)
    return NS(model=m, train=NS(type=train_type, dropout=0.0), data=NS(extra_pos=False, vari_data_len=False))
