from .forward import DecomposablePhysics, LinearPhysics, Physics  # noqa: F401
from .mri import MRI, MRIMixin, MultiCoilMRI  # noqa: F401
from .noise import GaussianNoise, NoiseModel, ZeroNoise  # noqa: F401
