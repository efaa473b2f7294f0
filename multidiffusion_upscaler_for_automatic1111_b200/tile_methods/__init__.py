from .abstractdiffusion import AbstractDiffusion
from .multidiffusion import MultiDiffusion
from .mixtureofdiffusers import MixtureOfDiffusers

__all__ = ["AbstractDiffusion", "MultiDiffusion", "MixtureOfDiffusers"]
