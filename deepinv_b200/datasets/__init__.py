"""Input side of the hot path: raw k-space readers that land data in the operators' layout."""
from .fastmri import FastMRISliceDataset, MRISliceTransform  # noqa: F401
