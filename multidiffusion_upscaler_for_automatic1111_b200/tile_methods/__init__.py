from .abstractdiffusion import AbstractDiffusion
from .multidiffusion import MultiDiffusion
from .mixtureofdiffusers import MixtureOfDiffusers
from .demofusion import DemoFusion

__all__ = ["AbstractDiffusion", "MultiDiffusion", "MixtureOfDiffusers", "DemoFusion"]
