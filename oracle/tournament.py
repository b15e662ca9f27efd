"""ORACLE (test infrastructure, never imported by the product path).

Restatement of the reference's tournament selection arithmetic, returning *positions* instead of
cloned agents.  Follows agilerl/hpo/tournament.py:41-119 (``_tournament`` draws from the GLOBAL
``np.random`` stream and takes the argmax over the *rank* vector; ``_elitism`` ranks by the mean
of the last ``eval_loop`` fitnesses with a double argsort)."""
from __future__ import annotations

import numpy as np


def select_positions(fitness_lists, indices, tournament_size, elitism, population_size, eval_loop):
    """-> (elite_pos, [(parent_pos, new_index), ...]) in the order the reference builds
    ``new_population`` (tournament.py:104-119)."""
    last_fitness = [np.mean(f[-eval_loop:]) for f in fitness_lists]
    rank = np.argsort(last_fitness).argsort()
    max_id = max(indices)
    elite_pos = int(np.argsort(rank)[-1])
    out = []
    n = population_size
    if elitism:
        out.append((elite_pos, indices[elite_pos]))
        n -= 1
    for _ in range(n):
        max_id += 1
        sel = np.random.randint(0, len(rank), size=tournament_size)
        vals = [rank[i] for i in sel]
        out.append((int(sel[np.argmax(vals)]), max_id))
    return elite_pos, out
