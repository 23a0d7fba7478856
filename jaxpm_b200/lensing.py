"""Light-cone density planes — the particle-facing part of /root/reference/jaxpm/lensing.py (density_plane :11-44).

One fused pass over the particles replaces the reference's mod / rescale / mask / cic_paint_2d chain (no xy and
weight arrays are materialised).  `convergence_Born` (ray tracing through the planes, lensing.py:47-90) is outside
the force-loop scope."""
import torch

from . import ops
from ._lib import as_f32, call, ptr, stream


def density_plane(positions, box_shape, center, width, plane_resolution, smoothing_sigma=None):
    """Extracts a density plane from the simulation (same arguments and normalisation as the reference)."""
    nx, ny, nz = box_shape
    pos = as_f32(positions).reshape(-1, 3)
    res = int(plane_resolution)
    plane = torch.zeros((res, res), dtype=torch.float32, device=pos.device)
    call("jpm_density_plane_f32", stream(), ptr(plane), ptr(pos), pos.shape[0], float(nx), float(center), float(width),
         res)
    # density normalisation, lensing.py:37-38
    norm = (nx / plane_resolution) * (ny / plane_resolution) * width
    plane = ops.axpby(1.0 / norm, plane, out=plane)
    if smoothing_sigma is not None:          # lensing.py:41-42
        from .utils import gaussian_smoothing
        plane = gaussian_smoothing(plane, smoothing_sigma)
    return plane
