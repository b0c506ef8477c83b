"""`demo/demo_tools.py` of the reference (imported by latent-space-interpolation-mnist.ipynb as `from demo_tools import *`):
resolves to the B200-native host mirror."""
from ladder_latent_data_distribution_modelling_b200.host.demo_tools import *  # noqa: F401,F403
from ladder_latent_data_distribution_modelling_b200.host.demo_tools import (  # noqa: F401
    MixtureDistribution, define_prior_distribution, generate_prior_embeddings, get_embeddings_from_val_set)
