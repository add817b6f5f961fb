"""Simulated experiment model `m2b` (reference experiment/models/m2b.py): see
hier_logistic.py for the definition shared by the logistic-regression family."""
from .hier_logistic import HierLogistic


class model(HierLogistic):
    family = 'm2b'
