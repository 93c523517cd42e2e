"""dpdfnet_b200 - B200-native batched streaming engine for DPDFNet's per-frame hot path.

Drop-in surface (same names as the reference package ``dpdfnet``, ``package/src/dpdfnet/__init__.py``):
``enhance``, ``available_models``, ``StreamEnhancer`` plus the sub-modules ``api``, ``stream``, ``audio``,
``onnx_backend`` and ``models``.  ``install_as_dpdfnet()`` registers the package under the name
``dpdfnet`` so unmodified callers (``import dpdfnet``) pick it up.
"""
__all__ = ["enhance", "enhance_batch", "available_models", "StreamEnhancer", "Engine", "install_as_dpdfnet"]


def __getattr__(name: str):
    if name in {"enhance", "enhance_batch", "available_models"}:
        from . import api
        return getattr(api, name)
    if name == "StreamEnhancer":
        from .stream import StreamEnhancer
        return StreamEnhancer
    if name == "Engine":
        from .engine import Engine
        return Engine
    raise AttributeError(f"module 'dpdfnet_b200' has no attribute '{name}'")


def install_as_dpdfnet() -> None:
    """Alias this package (and its mirror sub-modules) as ``dpdfnet`` in ``sys.modules``."""
    import importlib
    import sys
    pkg = sys.modules[__name__]
    sys.modules.setdefault("dpdfnet", pkg)
    for sub in ("api", "stream", "audio", "onnx_backend", "models"):
        sys.modules.setdefault(f"dpdfnet.{sub}", importlib.import_module(f"{__name__}.{sub}"))
