"""Simulated experiment model `m4b` (reference experiment/models/m4b.py): see
hier_logistic.py for the definition shared by the logistic-regression family."""
from .hier_logistic import HierLogistic


class model(HierLogistic):
    family = 'm4b'
