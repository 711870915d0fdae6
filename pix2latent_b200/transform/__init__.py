"""Transformation search (reference: pix2latent/transform/): jointly search an affine (and optionally
colour) transformation of the TARGET and the latent code. SURVEY.md §8f N2."""
from .spatial_transform import SpatialTransform
from .transform_optimizer import TransformBasinCMAOptimizer

__all__ = ["SpatialTransform", "TransformBasinCMAOptimizer"]
