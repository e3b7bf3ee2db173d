"""
hydromodel_b200 -- B200-native batched time-stepping core of the Berkeley 1-D
subsurface hydrological model (reference: vrettasm/HydroModel).

Layout
------
csrc/        CUDA kernels (sm_100a) and the C-ABI, built into lib/libhydrob200.so
capi.py      ctypes binding of the C-ABI (include/hydrob200.h)
pack.py      host precompute: JSON + forcing table -> flat device arrays
ensemble.py  EnsembleSimulation: many columns, sharded over GPUs
src/         drop-in mirror of the reference's Python objects
             (Simulation, RichardsPDE, SoilProperties, Porosity, ...)
synthetic.py deterministic synthetic forcing / parameter ensembles
"""
__version__ = "0.1.0"
