"""beer_b200 -- B200-native (sm_100a) Variational-Bayes E-step / M-step engine behind the
beer.inference / beer.models / beer.graph / beer.dists call surface.

    import beer_b200 as beer
    elbo = beer.evidence_lower_bound(model, X, datasize=N, inference_graph=graph)

Importing the package needs neither a GPU nor the shared library; calling any model method
does (there is no CPU fallback, see beer_b200/_lib.py)."""
__version__ = '0.2.0'

from . import dists, features, graph, vbi                                                    # noqa: F401
from .engine import BigramUnitWeights, CategoricalUnitWeights, EmissionParams, UnitWeights, Utterances, VBEngine, WeightGroup              # noqa: F401
from .dataset import Alignments, Dataset                                                      # noqa: F401
from .graph import CompiledGraph, Graph                                            # noqa: F401
from .inference import (EvidenceLowerBoundInstance, VBConjugateOptimizer, VBOptimizer,  # noqa: F401
                        evidence_lower_bound)
from .models import (HMM, Categorical, CategoricalSet, DiscreteLatentModel,        # noqa: F401
                     DynamicallyOrderedModelSet, JointModelSet, Mixture, MixtureSet, Model, ModelSet,
                     NormalSet, PhoneLoop, BigramPhoneLoop, SBCategorical, SBCategoricalHyperPrior)
from .vae import VAE                                                               # noqa: F401
from .parameters import BayesianParameter, ConjugateBayesianParameter             # noqa: F401
from .utils import logsumexp, onehot                                               # noqa: F401
