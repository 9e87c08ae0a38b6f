from pytorch_wavelets_b200.scatternet.layers import ScatLayer, ScatLayerj2  # noqa: F401
