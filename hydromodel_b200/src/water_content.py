"""Volumetric water-content bounds (value object).

Public surface of the reference's ``WaterContent``
(code/src/water_content.py:5-258): ``min, max, res`` must satisfy
``0 <= res < min < max <= 1`` at all times (setters roll back and raise
``ValueError`` otherwise); ``mid`` is their midpoint; ``wlt`` / ``flc`` are the
wilting point and field capacity suctions in cm (unconstrained).
"""
import numpy as np


class WaterContent:
    __slots__ = ("minimum", "maximum", "residual", "wilting", "field_cap")

    def __init__(self, minimum: float = 0.08, maximum: float = 0.30,
                 residual: float = 0.05, wilting: float = -1500.0,
                 field_cap: float = 340.0):
        self.minimum, self.maximum, self.residual = (
            WaterContent.inRange(v) for v in (minimum, maximum, residual))
        if not self._checkValues():
            raise ValueError(f" {type(self).__name__}: The volumetric water "
                             f"content input values are incorrect.")
        self.wilting = wilting
        self.field_cap = field_cap

    @staticmethod
    def inRange(new_value: float):
        """Clip a water content into [0, 1]."""
        return np.maximum(np.minimum(1.0, new_value), 0.0)

    def _checkValues(self):
        return 0.0 <= self.residual < self.minimum < self.maximum <= 1.0

    def _guarded_set(self, slot, label, value):
        previous = getattr(self, slot)
        setattr(self, slot, value)
        if not self._checkValues():
            setattr(self, slot, previous)
            raise ValueError(f" {type(self).__name__}: The new {label}: {value} is "
                             f"not consistent with the rest of the values.")

    min = property(lambda self: self.minimum,
                   lambda self, v: self._guarded_set("minimum", "minimum", v))
    max = property(lambda self: self.maximum,
                   lambda self, v: self._guarded_set("maximum", "maximum", v))
    res = property(lambda self: self.residual,
                   lambda self, v: self._guarded_set("residual", "residual", v))

    @property
    def mid(self):
        return 0.5 * (self.maximum + self.minimum)

    @property
    def wlt(self):
        return self.wilting

    @wlt.setter
    def wlt(self, value):
        self.wilting = value

    @property
    def flc(self):
        return self.field_cap

    @flc.setter
    def flc(self, value):
        self.field_cap = value

    def __str__(self):
        return (f" WaterContent Id({id(self)}): \n Minimum={self.minimum}, Maximum={self.maximum} \n"
                f" Residual={self.residual}, Wilting={self.wilting} \n Field capacity={self.field_cap}")
