"""module path of the reference's fault-injecting emulator (/root/reference/experiments/convergence/faulty_dmoe_emulator.py);
``get_non_expert_params`` here treats EmulatedFaultyDMoE layers as the expert-owning modules, like the reference's own copy"""
from functools import partial

from ...models.emulator import EmulatedFaultyDMoE, get_non_expert_params as _get_non_expert_params

get_non_expert_params = partial(_get_non_expert_params, dmoe_types=(EmulatedFaultyDMoE,))
