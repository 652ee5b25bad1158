from masic_b200.entropy_models import (EntropyBottleneck, EntropyModel, GaussianConditional,  # noqa: F401
                                       GaussianMixtureConditional, GaussianMixtureConditional_gf,
                                       pmf_to_quantized_cdf)

__all__ = ["EntropyModel", "EntropyBottleneck", "GaussianConditional", "GaussianMixtureConditional",
           "GaussianMixtureConditional_gf"]
