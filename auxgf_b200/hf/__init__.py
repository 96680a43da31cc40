from .model import ModelRHF, ModelUHF, ArraysRHF, ArraysUHF, Molecule, from_pyscf  # noqa: F401
