"""Result files.

The reference stores ``Simulation.output`` as root-level gzip-compressed HDF5
datasets through h5py (code/src/simulation.py:674-746).  h5py / libhdf5 are not
part of this image, so the I/O is layered: with h5py importable the files are
written and read exactly as the reference does; without it the same
name -> array mapping goes to a NumPy ``.npz`` archive next to the requested
``.h5`` name (documented gap: DESIGN.md, "out of scope / next").
"""
import os

import numpy as np


def _h5py():
    try:
        import h5py                      # noqa: WPS433 - optional dependency
        return h5py if hasattr(h5py, "File") else None
    except Exception:                    # noqa: BLE001
        return None


def save_datasets(path, data):
    """Write ``{name: ndarray}``; returns the path actually written."""
    h5 = _h5py()
    if h5 is not None:
        with h5.File(path, "w") as fh:
            for key, val in data.items():
                fh.create_dataset(key, data=val, shape=val.shape, compression="gzip")
        return str(path)
    alt = os.path.splitext(str(path))[0] + ".npz"
    np.savez_compressed(alt, **data)
    return alt


def load_datasets(path):
    """Inverse of :func:`save_datasets` (accepts the .h5 name for an .npz file)."""
    h5 = _h5py()
    path = str(path)
    if h5 is not None and os.path.exists(path) and not path.endswith(".npz"):
        with h5.File(path, "r") as fh:
            return {key: np.array(fh[key]) for key in fh}
    alt = path if path.endswith(".npz") else os.path.splitext(path)[0] + ".npz"
    with np.load(alt) as fh:
        return {key: fh[key] for key in fh.files}
