"""beer_b200 -- B200-native (sm_100a) Variational-Bayes E-step / M-step engine behind
the beer.inference / beer.models / beer.graph call surface."""
__version__ = '0.1.0'
