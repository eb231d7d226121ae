"""Legacy spellings of beer/vbi.py mapped onto the live API (beer_b200.inference), the way
SURVEY.md section 1 describes: beer/vbi.py is not imported by the reference's package and its
optimizers call parameter methods that no longer exist; the names are kept so that code
written against them keeps importing."""
from .inference import (EvidenceLowerBoundInstance, VBConjugateOptimizer, VBOptimizer,  # noqa: F401
                        add_acc_stats, evidence_lower_bound, scale_acc_stats)

EvidenceLowerBound = evidence_lower_bound                       # vbi.py:166-248
BayesianModelOptimizer = VBOptimizer                            # vbi.py:280-330
BayesianModelCoordinateAscentOptimizer = VBConjugateOptimizer   # vbi.py:333-380

__all__ = ['evidence_lower_bound', 'EvidenceLowerBound', 'EvidenceLowerBoundInstance', 'BayesianModelOptimizer',
           'BayesianModelCoordinateAscentOptimizer', 'VBConjugateOptimizer', 'VBOptimizer']
