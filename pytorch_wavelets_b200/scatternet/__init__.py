from pytorch_wavelets_b200.scatternet.layers import ScatLayer  # noqa: F401
