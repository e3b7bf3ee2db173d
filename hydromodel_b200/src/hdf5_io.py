"""Result files.

The reference stores ``Simulation.output`` as root-level gzip-compressed HDF5
datasets through h5py (code/src/simulation.py:674-746).  With h5py importable the
files are written and read exactly as the reference does; without it (this image)
the same datasets go through the built-in writer / reader of ``hdf5_min`` (HDF5
superblock 0, version-1 object headers, one deflate-compressed chunk per dataset:
the structures h5py itself writes with its default format bounds).  ``.npz``
archives written by earlier versions of this package are still read.
"""
import os

import numpy as np

from . import hdf5_min


def _h5py():
    try:
        import h5py                      # noqa: WPS433 - optional dependency
        return h5py if hasattr(h5py, "File") else None
    except Exception:                    # noqa: BLE001
        return None


def save_datasets(path, data):
    """Write ``{name: ndarray}`` to the HDF5 file ``path``; returns the path."""
    h5 = _h5py()
    if h5 is not None:
        with h5.File(path, "w") as fh:
            for key, val in data.items():
                fh.create_dataset(key, data=val, shape=val.shape, compression="gzip")
        return str(path)
    # gzip level 4 is h5py's default for compression='gzip' (simulation.py:703)
    return hdf5_min.write_file(str(path), {k: np.asarray(v) for k, v in data.items()},
                               compression_level=4)


def load_datasets(path):
    """Inverse of :func:`save_datasets` (the reference's ``loadResults``)."""
    h5 = _h5py()
    path = str(path)
    if os.path.exists(path) and not path.endswith(".npz"):
        if h5 is not None:
            with h5.File(path, "r") as fh:
                return {key: np.array(fh[key]) for key in fh}
        return hdf5_min.read_file(path)
    alt = path if path.endswith(".npz") else os.path.splitext(path)[0] + ".npz"
    with np.load(alt) as fh:
        return {key: fh[key] for key in fh.files}
