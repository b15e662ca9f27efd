"""Stand-ins for third-party packages the reference's off-policy path imports but the B200
build image does not ship (``tensordict``, ``gymnasium``).  The real package wins when present."""
try:  # pragma: no cover - depends on the image
    from tensordict import TensorDict, TensorDictBase, is_tensor_collection, tensorclass  # type: ignore
    HAVE_TENSORDICT = True
except Exception:  # noqa: BLE001
    from .tensordict import TensorDict, TensorDictBase, is_tensor_collection, tensorclass
    HAVE_TENSORDICT = False

try:  # pragma: no cover - depends on the image
    from gymnasium import spaces  # type: ignore
    HAVE_GYMNASIUM = True
except Exception:  # noqa: BLE001
    from . import spaces
    HAVE_GYMNASIUM = False

__all__ = ["TensorDict", "TensorDictBase", "is_tensor_collection", "tensorclass", "spaces"]
