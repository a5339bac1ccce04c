"""On-disk DEM format (the reference has no map format at all -- only PCD submap dumps,
mbes_mapper/src/mbes_receptor.cpp:106, and a landmark YAML service, map_provider_node.py:35-48).

A DEM file is a NumPy ``.npz`` with

    z        (H, W) float32 vertex heights [m], row i at y0 + i*res, column j at x0 + j*res     -- or --
    zq       (H, W) uint16 quantised heights,  z = z_min + zq * z_step   (optional compact form, 2 B / vertex)
    x0, y0, res        float64 scalars
    z_min, z_step      float64 scalars (with zq only)
    frame              str, the map frame name (default "map")

uint16 heights halve the file and the upload; they are DEQUANTISED ON LOAD to float32, and that float32 array is
what both the GPU library and the oracle see -- so parity holds bit for bit on the quantised map, and the
quantisation error (z_step / 2, e.g. 0.5 mm for 65 m of relief) is a property of the MAP, not of the kernels.
"""
import numpy as np


def quantise(z, z_step=None):
    """(zq uint16, z_min, z_step): z_step defaults to the finest grid that covers the relief."""
    z = np.asarray(z, np.float32)
    z_min = float(z.min())
    span = float(z.max()) - z_min
    if z_step is None:
        z_step = span / 65535.0 if span > 0 else 1.0
    zq = np.rint((z.astype(np.float64) - z_min) / z_step)
    if zq.max() > 65535:
        raise ValueError("z_step %g too fine for %g m of relief" % (z_step, span))
    return zq.astype(np.uint16), z_min, float(z_step)


def dequantise(zq, z_min, z_step):
    """The float32 heights every consumer sees: float32(z_min + zq * z_step), evaluated in float64."""
    return (np.float64(z_min) + zq.astype(np.float64) * np.float64(z_step)).astype(np.float32)


def save_dem(path, dem, quantised=False, z_step=None, frame="map"):
    """``dem``: {"z", "x0", "y0", "res"} (the dict ``ParticleFilter.set_dem`` takes)."""
    meta = dict(x0=np.float64(dem["x0"]), y0=np.float64(dem["y0"]), res=np.float64(dem["res"]), frame=np.str_(frame))
    if quantised:
        zq, z_min, step = quantise(dem["z"], z_step)
        np.savez_compressed(path, zq=zq, z_min=np.float64(z_min), z_step=np.float64(step), **meta)
    else:
        np.savez_compressed(path, z=np.asarray(dem["z"], np.float32), **meta)


def load_dem(path):
    """-> {"z" float32 (H, W), "x0", "y0", "res", "frame"} ready for ``ParticleFilter.set_dem``."""
    with np.load(path, allow_pickle=False) as f:
        if "zq" in f.files:
            z = dequantise(f["zq"], float(f["z_min"]), float(f["z_step"]))
        elif "z" in f.files:
            z = np.ascontiguousarray(f["z"], np.float32)
        else:
            raise ValueError("%s: neither 'z' nor 'zq'" % path)
        if z.ndim != 2 or min(z.shape) < 2:
            raise ValueError("%s: heights must be (H, W) with H, W >= 2" % path)
        if not np.isfinite(z).all():
            raise ValueError("%s: non-finite heights" % path)
        res = float(f["res"])
        if not res > 0:
            raise ValueError("%s: res must be positive" % path)
        return {"z": z, "x0": float(f["x0"]), "y0": float(f["y0"]), "res": res,
                "frame": str(f["frame"]) if "frame" in f.files else "map"}
