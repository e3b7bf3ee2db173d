"""Saturated hydraulic conductivities and noise-model parameters (value object).

Public surface of the reference's ``HydraulicConductivity``
(code/src/hydraulic_conductivity.py:3-240).  Units: cm per 0.5 h.  The three
saturated values must be strictly positive, ``sigma_noise`` and
``lambda_exponent`` non-negative (the constructor clamps those two at zero,
the setters raise ``ValueError``).
"""
import numpy as np

_LABELS = {"sat_soil": "saturated value of the Soil layer",
           "sat_saprolite": "saturated value of the Saprolite layer",
           "sat_fresh_bedrock": "saturated value of the Fresh Bedrock layer",
           "sigma_noise": "sigma amplitude value of the noise model",
           "lambda_exponent": "lambda exponent value of the noise model"}


def _field(name, strict):
    slot = "param_" + name

    def setter(self, value):
        if (value > 0.0) if strict else (value >= 0.0):
            setattr(self, slot, value)
        else:
            need = "strictly positive" if strict else "non-negative"
            raise ValueError(f" {type(self).__name__}: The {_LABELS[name]}: "
                             f"{value} should be {need}.")

    return property(lambda self: getattr(self, slot), setter, doc=_LABELS[name])


class HydraulicConductivity:
    __slots__ = ("param_sat_soil", "param_sat_saprolite", "param_sigma_noise",
                 "param_sat_fresh_bedrock", "param_lambda_exponent")

    sat_soil = _field("sat_soil", True)
    sat_saprolite = _field("sat_saprolite", True)
    sat_fresh_bedrock = _field("sat_fresh_bedrock", True)
    sigma_noise = _field("sigma_noise", False)
    lambda_exponent = _field("lambda_exponent", False)

    def __init__(self, sat_soil: float = 8.5, sat_saprolite: float = 3.2,
                 sat_fresh_bedrock: float = 0.1, sigma_noise: float = 2.0,
                 lambda_exponent: float = 1.0):
        self.sat_soil = sat_soil
        self.sat_saprolite = sat_saprolite
        self.sat_fresh_bedrock = sat_fresh_bedrock
        self.param_sigma_noise = np.maximum(sigma_noise, 0.0)
        self.param_lambda_exponent = np.maximum(lambda_exponent, 0.0)

    def __str__(self):
        return (f" HydraulicConductivity Id({id(self)}): \n Sat-Soil={self.sat_soil} \n"
                f" Sat-Saprolite={self.sat_saprolite} \n Sat-Fresh-Bedrock={self.sat_fresh_bedrock} \n"
                f" Sigma={self.sigma_noise} \n Lambda={self.lambda_exponent}")
