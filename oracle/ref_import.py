"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference (Ali2500/Video-LLaVA-Seg,
vendored SAM 2.1 under /root/reference/sam2) on CPU so that golden vectors can be generated
and the oracle restatement (oracle/sam2_path.py) can be pinned against it.

/root/reference exists only in the build container; nothing in `-m gpu` tests, smoke() or
bench.py may import this module.  Recipe follows SURVEY.md Appendix A:
  * `sam2/__init__.py` needs hydra            -> register a namespace stub for `sam2`
  * `backbones/hieradet.py:14` needs iopath    -> stub `iopath.common.file_io.g_pathmgr`
  * YAML `_target_` instantiation (build_sam.py:79-118) without hydra/omegaconf
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("VLS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "sam2", "modeling"))


def _install_stubs():
    if "sam2" not in sys.modules or not hasattr(sys.modules["sam2"], "__path__"):
        pkg = types.ModuleType("sam2")
        pkg.__path__ = [os.path.join(REF_ROOT, "sam2")]
        sys.modules["sam2"] = pkg
    for n in ("iopath", "iopath.common", "iopath.common.file_io"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["iopath.common.file_io"].g_pathmgr = None


def _inst(node):
    if isinstance(node, dict):
        kw = {k: _inst(v) for k, v in node.items() if k != "_target_"}
        if "_target_" in node:
            mod, cls = node["_target_"].rsplit(".", 1)
            return getattr(importlib.import_module(mod), cls)(**kw)
        return kw
    if isinstance(node, list):
        return [_inst(x) for x in node]
    if isinstance(node, str):
        try:
            return float(node)  # PyYAML reads `1e-6` as str (sam2.1_hiera_b+.yaml:80)
        except ValueError:
            return node
    return node


def load_cfg(variant: str = "b+"):
    import yaml

    path = os.path.join(REF_ROOT, "sam2", "configs", "sam2.1", f"sam2.1_hiera_{variant}.yaml")
    with open(path) as f:
        return yaml.safe_load(f)["model"]


def build_video_predictor(variant: str = "b+", seed: int = 0, with_image_encoder: bool = True):
    """Reference SAM2VideoPredictor with the build_sam.py:88-102 overrides applied by hand."""
    import torch

    _install_stubs()
    cfg = load_cfg(variant)
    cfg["_target_"] = "sam2.sam2_video_predictor.SAM2VideoPredictor"  # build_sam.py:89
    cfg["binarize_mask_from_pts_for_mem_enc"] = True  # build_sam.py:99
    cfg["fill_hole_area"] = 8  # build_sam.py:101
    cfg["sam_mask_decoder_extra_args"] = dict(  # build_sam.py:95-97 (inert in this fork)
        dynamic_multimask_via_stability=True,
        dynamic_multimask_stability_delta=0.05,
        dynamic_multimask_stability_thresh=0.98,
    )
    torch.manual_seed(seed)
    model = _inst(cfg).eval()
    if not with_image_encoder:
        model.image_encoder = None
    return model


def ref_modules():
    """Return the reference module namespaces (after stubbing)."""
    _install_stubs()
    import sam2.modeling.memory_attention as ma
    import sam2.modeling.memory_encoder as me
    import sam2.modeling.sam.mask_decoder as md
    import sam2.modeling.sam.transformer as tr
    import sam2.modeling.position_encoding as pe

    return types.SimpleNamespace(memory_attention=ma, memory_encoder=me, mask_decoder=md, transformer=tr,
                                 position_encoding=pe)
