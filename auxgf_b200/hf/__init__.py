from .model import ModelRHF, ModelUHF, Molecule  # noqa: F401
