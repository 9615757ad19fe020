from .diffusion import DDRM, DiffPIR  # noqa: F401
