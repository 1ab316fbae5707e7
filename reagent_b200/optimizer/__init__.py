from .fused_adam import FusedAdam  # noqa: F401
from .soft_update import SoftUpdate  # noqa: F401
from .union import Adam, Optimizer__Union  # noqa: F401

__all__ = ["Optimizer__Union", "SoftUpdate", "FusedAdam", "Adam"]
