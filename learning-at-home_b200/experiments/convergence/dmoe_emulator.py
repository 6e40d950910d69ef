"""module path of the reference emulator (/root/reference/experiments/convergence/dmoe_emulator.py): the notebooks and user
code do ``from dmoe_emulator import EmulatedDMoE, get_non_expert_params``; the implementation lives in ``models/emulator.py``"""
from ...models.emulator import EmulatedDMoE, get_non_expert_params  # noqa: F401
