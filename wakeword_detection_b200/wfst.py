"""Superframe smoothing of keyword posteriors - the decoding the reference builds as a weighted FST and solves with
pynini's shortest path (reference: wwdetect/wfst.py:17-71).

The lattice has one state per (timepoint, label) with labels ('other', 'wakeword'); the arc into label p at time t costs
-log P[t, p], minus 1 when it stays in the label it comes from; the start arcs cost -log(1/2) - log P[0, p].  The best
path of such a trellis is a Viterbi recursion - no FST library needed (pynini / graphviz are not installable here)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

LABELS = ("other", "wakeword")
STAY_BONUS = ((1.0, 1.0), (1.0, 1.0))      # test_transition_matrix (wfst.py:24-25)


def best_path(posterior_probs: Sequence[Sequence[float]]) -> Tuple[List[int], float]:
    """Label index per timepoint of the cheapest path and its total cost."""
    obs = -np.log(np.asarray(posterior_probs, np.float64))
    T, P = obs.shape
    cost = -np.log(1.0 / P) + obs[0]
    back = np.zeros((T, P), np.int64)
    for t in range(1, T):
        new = np.empty(P)
        for p_to in range(P):
            cand = [cost[p_from] + obs[t, p_to] - (STAY_BONUS[p_to][p_from] if p_to == p_from else 0.0) for p_from in range(P)]
            back[t, p_to] = int(np.argmin(cand))
            new[p_to] = cand[back[t, p_to]]
        cost = new
    last = int(np.argmin(cost))
    path = [last]
    for t in range(T - 1, 0, -1):
        path.append(int(back[t, path[-1]]))
    return path[::-1], float(cost[last])


def smooth(posterior_probs) -> str:
    """wfst.py:17-71: the smoothed label string ('other other wakeword ...'), which the reference prints and returns as
    the shortest-path FST."""
    path, _ = best_path(posterior_probs)
    return " ".join(LABELS[p] for p in path)
