"""B200-native implementation of the EP inner loop of gelman/ep-stan, behind the
reference's own package layout: ``epstan.method`` (Master, Worker) and
``epstan.util``."""
