from .diffusion import DDRM  # noqa: F401
