"""hmc_ft_b200 — B200 (sm_100a) implementation of the 2D U(1) HMC / FT-HMC hot path of
Greyyy-HJC/hmc_ft behind the reference's own class surface.  See DESIGN.md / INTEGRATION.md."""
from . import _lib
from ._lib import U1Error
from .hmc_u1 import HMC_U1

__all__ = ["HMC_U1", "U1Error", "_lib"]

try:  # FT classes (need the same library)
    from .field_trans import FieldTransformation  # noqa: F401
    from .hmc_u1_ft import HMC_U1_FT  # noqa: F401
    __all__ += ["FieldTransformation", "HMC_U1_FT"]
except ImportError:  # pragma: no cover - only while the FT host files are absent
    pass
