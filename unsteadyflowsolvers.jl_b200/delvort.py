"""Vortex-count control definitions, src/delVort.jl:1-9."""
from dataclasses import dataclass


class delVortDef:
    pass


@dataclass
class delNone(delVortDef):
    """delVort.jl:3"""


@dataclass
class delSpalart(delVortDef):
    """delVort.jl:5-9; defaults from the docstring at calcs.jl:524"""
    limit: int = 500
    dist: float = 10.0
    tol: float = 1e-6
