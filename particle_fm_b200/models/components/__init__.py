from .epic import EPiC_encoder, EPiC_layer  # noqa: F401
