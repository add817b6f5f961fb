"""Simulated experiment model `m3b` (reference experiment/models/m3b.py): see
hier_logistic.py for the definition shared by the logistic-regression family."""
from .hier_logistic import HierLogistic


class model(HierLogistic):
    family = 'm3b'
