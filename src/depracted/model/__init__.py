"""`from src.depracted.model import SpatialDROW, Prototype` works here (the reference package __init__ is empty: SURVEY.md D1)."""
from planar_optical_flow_b200.model.dr_spaam import DROW, SpatialDROW, _SpatialAttention  # noqa: F401
from planar_optical_flow_b200.model.prototype import Prototype  # noqa: F401
