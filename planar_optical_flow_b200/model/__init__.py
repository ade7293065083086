from .dr_spaam import DROW, SpatialDROW, _SpatialAttention  # noqa: F401
