"""Simulated experiment model `m5b` (reference experiment/models/m5b.py): see
hier_logistic.py for the definition shared by the logistic-regression family."""
from .hier_logistic import HierLogistic


class model(HierLogistic):
    family = 'm5b'
