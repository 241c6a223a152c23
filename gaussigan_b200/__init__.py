"""gaussigan_b200 -- B200-native (sm_100a) drop-in for GaussiGAN's splat-to-silhouette renderer.

Only the reference's data-parallel hot path lives here: the renderer
``get_landmarks`` / ``get_gaussian_maps_2d`` and its backward, as hand-written CUDA kernels behind a
C ABI (``include/gaussigan.h``), with this package as the host-side mirror of the reference's Python
interface.  Importing it loads ``lib/libgaussigan_sm100.so`` and fails loudly if it is missing.
"""
from ._lib import GaussiganError, lib as _capi, set_fast_mask, set_pdl  # noqa: F401  (import = load the shared library)
from .renderer import (REFERENCE_SIZES, get_gaussian_maps_2d, get_landmarks, get_landmarks_infer,
                       get_landmarks_into, get_landmarks_views, get_landmarks_with_mask, get_silhouette_2d, project, render_silhouette)

from .consumers import (colorize_gaussians, colorize_landmark_maps, gm_cycle_loss, mask_recon_loss_bis,  # noqa: E402
                        turntable)
from .geometry import deform, sigma_from_frame, svd3  # noqa: E402
from .interchange import decoder_feed, load_reference_gaussians, save_reference_gaussians  # noqa: E402

__all__ = ["set_fast_mask", "set_pdl", "svd3", "deform", "sigma_from_frame", "colorize_gaussians", "colorize_landmark_maps", "gm_cycle_loss", "mask_recon_loss_bis", "turntable",
           "decoder_feed", "load_reference_gaussians", "save_reference_gaussians",
           "REFERENCE_SIZES", "GaussiganError", "get_gaussian_maps_2d", "get_landmarks", "get_landmarks_infer",
           "get_landmarks_into", "get_landmarks_views", "get_landmarks_with_mask", "get_silhouette_2d", "project", "render_silhouette"]
__version__ = "0.2.0"
