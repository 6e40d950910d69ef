"""
Single-process emulators of a DMoE layer (convergence experiments, BASELINE config #1):

* ``EmulatedDMoE``       — /root/reference/experiments/convergence/dmoe_emulator.py:6-82
* ``EmulatedFaultyDMoE`` — /root/reference/experiments/convergence/faulty_dmoe_emulator.py:6-86 (Bernoulli expert failures)

Same constructor, state_dict layout (``expert_keys``, ``gating_pre_normalize.*``, ``experts.{i}.*``,
``expert_inputs_since_update``, ``expert_steps_since_first_input``) and update rule: every expert owns an optimizer that
is stepped automatically once the expert has seen >= update_every_inputs inputs or >= update_every_steps steps since
its first input.  The implementation is different: tokens are grouped per expert and every chosen expert runs ONCE on
the batch of its tokens (the reference calls each expert on single vectors inside a per-sample Python loop).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class EmulatedDMoE(nn.Module):
    failure_rate = 0.0

    def __init__(self, in_features, num_experts, num_active, update_every_inputs, update_every_steps, Expert, Optimizer):
        super().__init__()
        self.gating_pre_normalize = nn.LayerNorm(in_features)
        self.expert_keys = nn.Parameter(torch.randn(in_features, num_experts))
        self.experts = nn.ModuleList([Expert(in_features) for _ in range(num_experts)])
        self.expert_optimizers = {expert: Optimizer(expert.parameters()) for expert in self.experts}
        self.register_buffer("expert_inputs_since_update", torch.zeros(num_experts, dtype=torch.int64))
        self.register_buffer("expert_steps_since_first_input", torch.zeros(num_experts, dtype=torch.int64))
        self.num_active = num_active
        self.update_every_inputs, self.update_every_steps = update_every_inputs, update_every_steps

    # ------------------------------------------------------------------ gating
    def gating_logits(self, input):
        # NOTE: normalisation is along the expert axis (dim=-1 of [in_features, num_experts]), as in the reference
        logits = self.gating_pre_normalize(input) @ F.normalize(self.expert_keys, dim=-1)
        if self.failure_rate:
            failed = torch.rand_like(logits) < self.failure_rate
            logits = logits.masked_fill(failed, float("-inf"))
        return logits

    def forward(self, input):
        assert input.dim() == 2
        if self.training:
            self.maybe_update_experts()

        logits = self.gating_logits(input)
        top_logits, chosen_ids = torch.topk(logits, self.num_active, dim=-1, sorted=True)
        # failed experts have -inf logits => zero weight; a sample whose experts ALL failed gets a zero output
        # (the reference produces NaN there: softmax over an all -inf row)
        weights = F.softmax(top_logits, dim=-1).nan_to_num(0.0)

        # group (sample, slot) pairs by expert; run each used expert once on all of its tokens
        flat_ids = chosen_ids.reshape(-1)
        order = torch.argsort(flat_ids, stable=True)
        sorted_ids = flat_ids[order]
        token_of = torch.div(order, self.num_active, rounding_mode="floor")
        used, counts = torch.unique_consecutive(sorted_ids, return_counts=True)
        expert_out = torch.empty(flat_ids.numel(), input.shape[1], dtype=input.dtype, device=input.device)
        start = 0
        for expert_id, count in zip(used.tolist(), counts.tolist()):
            rows = order[start: start + count]
            expert_out = expert_out.index_copy(0, rows, self.experts[expert_id](input[token_of[start: start + count]]))
            start += count
        expert_out = expert_out.view(input.shape[0], self.num_active, -1)
        output = torch.einsum("bkd,bk->bd", expert_out, weights)

        if self.training:
            with torch.no_grad():
                self.expert_inputs_since_update.scatter_add_(0, flat_ids, torch.ones_like(flat_ids))
                self.expert_steps_since_first_input += (self.expert_inputs_since_update > 0).to(torch.int64)
        return output

    # ------------------------------------------------------------------ asynchronous expert updates
    def maybe_update_experts(self):
        due = (self.expert_inputs_since_update >= self.update_every_inputs) | \
              (self.expert_steps_since_first_input >= self.update_every_steps)
        for i in torch.nonzero(due).flatten().tolist():
            optimizer = self.expert_optimizers[self.experts[i]]
            optimizer.step()
            optimizer.zero_grad()
        self.expert_inputs_since_update[due] = 0
        self.expert_steps_since_first_input[due] = 0


class EmulatedFaultyDMoE(EmulatedDMoE):
    """EmulatedDMoE + fault injection: each (sample, expert) pair fails independently with prob. failure_rate; failed
    experts are excluded from the top-k and the softmax renormalises over the survivors."""

    def __init__(self, in_features, num_experts, num_active, update_every_inputs, update_every_steps, failure_rate,
                 Expert, Optimizer):
        super().__init__(in_features, num_experts, num_active, update_every_inputs, update_every_steps, Expert, Optimizer)
        self.failure_rate = failure_rate


def get_non_expert_params(model, dmoe_types=(EmulatedDMoE,)):
    """Parameters the TRAINER optimises: everything outside DMoE layers (experts optimise themselves; like the
    reference this also leaves the emulator's gate — expert_keys / gating_pre_normalize — out of the trainer)."""
    owned = {id(p) for m in model.modules() if isinstance(m, dmoe_types) for p in m.parameters()}
    return [p for p in model.parameters() if id(p) not in owned]
