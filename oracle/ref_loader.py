"""Import the reference's OWN hot-path modules from /root/reference (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Used by oracle/make_golden.py to
generate the committed fixtures and by tests (skipped when /root/reference is
absent, e.g. on the GPU box) to validate oracle/restatement.py.  No reference
source is copied: modules are imported from where they lie.

Recipe: pytorch3d -> oracle.p3d_shim; viz / optimisation
dependencies that the hot path never calls -> MagicMock; namespace stubs for
``lib_shape_prior.core...`` so that ``lib_shape_prior/core/__init__.py`` (which
drags in the training solver, compiled Cython extensions and ``torch._six``) is
never executed.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from unittest import mock

import torch

from . import p3d_shim

REF_ROOT = os.environ.get("LS_REFERENCE_ROOT", "/root/reference")
CKPT = "weights/checkpoint/LivingScenes_latest.pt"
FIELD_CFG = "weights/files_backup/model_config.yaml"

_state: dict = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model_utils.py"))


def checkpoint_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, CKPT))


def _ns(name: str, path: str) -> None:
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m


def load():
    """Return a namespace with the reference modules: model_utils, matcher_new,
    pose_estimation, vec_dgcnn_atten, deepsdf_decoder."""
    if "mods" in _state:
        return _state["mods"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    p3d_shim.install()
    for name in ["pycg", "matplotlib", "matplotlib.pyplot", "trimesh", "point_cloud_utils",
                 "torchlie", "geomloss", "roma", "coloredlogs", "open3d"]:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = mock.MagicMock()
    lsp = os.path.join(REF_ROOT, "lib_shape_prior")
    _ns("lib_shape_prior", lsp)
    _ns("lib_shape_prior.core", os.path.join(lsp, "core"))
    _ns("lib_shape_prior.core.models", os.path.join(lsp, "core", "models"))
    _ns("lib_shape_prior.core.models.utils", os.path.join(lsp, "core", "models", "utils"))
    _ns("lib_shape_prior.core.models.utils.occnet_utils",
        os.path.join(lsp, "core", "models", "utils", "occnet_utils"))
    sys.modules["lib_shape_prior.core.models.utils.occnet_utils.mesh_extractor2"] = mock.MagicMock()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    mods = types.SimpleNamespace()
    mods.model_utils = importlib.import_module("model_utils")
    mods.matcher_new = importlib.import_module("lib_more.matcher_new")
    mods.pose_estimation = importlib.import_module("lib_more.pose_estimation")
    mods.vec_dgcnn_atten = importlib.import_module("lib_shape_prior.core.lib.vec_sim3.vec_dgcnn_atten")
    mods.deepsdf_decoder = importlib.import_module("lib_shape_prior.core.lib.implicit_func.deepsdf_decoder")
    _state["mods"] = mods
    return mods


def shape_prior(state_dict=None):
    """Build the reference's ``Shape_Prior`` (use_double=False, as configs/room4cates.yaml:15).

    With ``state_dict=None`` the shipped checkpoint is loaded exactly as the
    reference does (model_utils.py:118-128, plus map_location='cpu').  Otherwise
    ``state_dict`` (keys ``encoder.*`` / ``decoder.*``) replaces the weights.
    """
    mods = load()
    real_load = torch.load

    def cpu_load(f, *a, **k):
        k.setdefault("map_location", "cpu")
        k.setdefault("weights_only", True)
        return real_load(f, *a, **k)

    cfg = {"working_dir": REF_ROOT, "field_cfg": FIELD_CFG, "field_pt": CKPT}
    with mock.patch.object(torch, "load", cpu_load):
        sp = mods.model_utils.Shape_Prior(cfg, "chair", use_double=False)
    sp.eval()
    if state_dict is not None:
        enc = {k[len("encoder."):]: v for k, v in state_dict.items() if k.startswith("encoder.")}
        dec = {k[len("decoder."):]: v for k, v in state_dict.items() if k.startswith("decoder.")}
        sp.encoder.load_state_dict(enc, strict=True)
        sp.decoder.F.load_state_dict(dec, strict=True)
    return sp


def shipped_state_dict() -> dict:
    """The shipped checkpoint's encoder/decoder tensors under ``encoder.*`` / ``decoder.*`` keys."""
    ck = torch.load(os.path.join(REF_ROOT, CKPT), map_location="cpu", weights_only=True)
    sd = ck["model_state_dict"]
    out = {}
    for k, v in sd.items():
        parts = k.split(".")
        if parts[0] == "network_dict" and parts[1] in ("encoder", "decoder"):
            out[".".join(parts[1:])] = v.detach().clone().float().contiguous()
    return out
