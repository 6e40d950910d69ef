"""
Host-side model of the hot-expert shadow selection that ``layout_exchange_kernel`` (csrc/moe.cu) runs on every GPU.

Used (a) as the executable specification of the kernel — ``tools/multi_gpu_check.py`` compares the kernel's choice with
this function on the real count tables — and (b) offline, to predict the load balance of a routing histogram
(``tools/routing_probe.py``).  The reference has no counterpart: a hot expert simply queues requests in its TaskPool
(/root/reference/lib/runtime/task_pool.py:105-125).
"""
from typing import List, Sequence, Tuple


def rank_loads(counts: Sequence[Sequence[int]], E_loc: int, shadowed: Sequence[int] = ()) -> List[int]:
    """rows every rank processes: all rows of the experts it owns that are not shadowed + its OWN rows of shadowed ones"""
    world, E = len(counts), len(counts[0])
    sh = set(shadowed)
    loads = [0] * world
    for e in range(E):
        if e in sh:
            for r in range(world):
                loads[r] += counts[r][e]
        else:
            loads[e // E_loc] += sum(counts[r][e] for r in range(world))
    return loads


def shadow_plan(counts: Sequence[Sequence[int]], E_loc: int, max_shadow: int, tol: float = 1.1,
                min_rows: int = 1) -> Tuple[List[int], List[int]]:
    """
    :param counts: [world][E] rows routed by every rank to every expert (expert e lives on rank e // E_loc)
    :returns: (shadowed experts in selection order, resulting rank loads)

    Greedy, deterministic (ties -> smaller index): while the most loaded rank exceeds ``tol`` x the mean load, its largest
    not-yet-shadowed expert is shadowed, unless that expert has fewer than ``min_rows`` rows in total.
    The greedy step always relieves the owner; it is not guaranteed to lower the maximum for adversarial count matrices (a
    nearly-as-loaded sender that contributed most of the expert's rows gets them back) — with real routing every rank
    contributes ~1/world of an expert's rows (tests/test_properties.py).
    """
    world, E = len(counts), len(counts[0])
    tot = [sum(counts[r][e] for r in range(world)) for e in range(E)]
    loads = rank_loads(counts, E_loc)
    shadowed: List[int] = []
    if world == 1:
        return shadowed, loads
    for _ in range(max_shadow):
        total = sum(loads)
        rmax = max(range(world), key=lambda r: (loads[r], -r))
        if loads[rmax] * world <= tol * total:
            break
        cand = [e for e in range(rmax * E_loc, (rmax + 1) * E_loc) if e not in shadowed]
        if not cand:
            break
        best = max(cand, key=lambda e: (tot[e], -e))
        if tot[best] < min_rows:
            break
        shadowed.append(best)
        loads[rmax] -= tot[best]
        for r in range(world):
            loads[r] += counts[r][best]
    return shadowed, loads
