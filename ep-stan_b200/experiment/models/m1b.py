"""Simulated experiment model `m1b` (reference experiment/models/m1b.py): see
hier_logistic.py for the definition shared by the logistic-regression family."""
from .hier_logistic import HierLogistic


class model(HierLogistic):
    family = 'm1b'
