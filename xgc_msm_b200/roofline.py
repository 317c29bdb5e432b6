"""Algorithmic HBM bytes of each weekly kernel (DESIGN.md §5 states the same numbers).

"Algorithmic" = bytes the kernel's work item must move given the packed layout of xgc_device.cuh, counted once per
reference: a whole plane element when any field of it is needed, gathers counted per endpoint.  These are the
numerators of roofline.achieved in bench.py; ncu's dram__bytes_read+write per launch is the cross-check (traffic).

Plane element sizes: core 16 B, aux 16 B, each clock plane 8 B, edge record 16 B.
Population fractions come from the week's own tallies (HIV prevalence, diagnosed fraction, GC prevalence).
"""
from __future__ import annotations


def kernel_bytes(stats: dict) -> dict:
    """stats: n (living agents, all replicates), f_hiv, f_dx, f_gc (fractions of n), edges (all layers, all replicates),
    f_disc (fraction of edges with any discordance; only used for k_edge2's optional aux gathers).
    Returns {kernel: algorithmic bytes per launch}."""
    n, e = stats["n"], stats["edges"]
    f_hiv, f_dx, f_gc = stats["f_hiv"], stats["f_dx"], stats["f_gc"]
    f_disc = stats.get("f_disc", 0.6)        # discordant (edge, act type) pairs per edge (measured: queue lengths / edges)
    return {
        # core R + aux R (age for aging/mortality) + tck R (last.neg.test, undiagnosed) (+ 4 B queue entry per HIV-positive)
        "k_agent1a": n * (16 + 16 + 8 * (1 - f_dx) + f_hiv * 4),
        # queued HIV-positive agents: item + core + aux (vl) + hck_a R + hck_b R+W
        "k_agent1b": n * f_hiv * (4 + 16 + 16 + 8 + 8 + 8),
        # prep + classification: core R + tck R (+ 4 B queue entry for agents with GC work)
        "k_agent2a": n * (16 + 8 + f_gc * 4),
        # queued agents (infected or visiting): item + core + gck_a + gck_b
        "k_agent2b": n * f_gc * (4 + 16 + 8 + 8),
        # tally pass: core R (+ 4 B queue entry per agent with a new infection)
        "k_agent3a": n * (16 + stats.get("f_new", 0.0) * 4),
        # queued agents: item + core R+W (8 B of it) + gck_a / gck_b R+W
        "k_agent3b": n * stats.get("f_new", 0.0) * (4 + 16 + 8 + 4 * 8),
        # edge record R + acts W(4) + 2 endpoints x (core + aux)
        "k_edge1": e * (16 + 4 + 2 * 32),
        # classify: edge record R + 2 endpoints x core (+ 4 B queue entry per discordant (edge, type) pair)
        "k_edge2a": e * (16 + 2 * 16 + f_disc * 4),
        # dense queue: per pair 4 B item + edge record + 2 x core (+ aux for anal pairs, ~1/4 of the pairs)
        "k_edge2b_anal": e * f_disc * 0.25 * (4 + 16 + 2 * 16 + 2 * 16),
        "k_edge2b_oral_kiss_rim": e * f_disc * 0.75 * (4 + 16 + 2 * 16),
        # in-place stable compaction on the alive bitmap (1 bit per endpoint, cache-resident) with the dissolution flag pre-drawn in the
        # record: record R, and W only of the records that move (everything behind the first removed tie of a layer: about half);
        # ncu: 93 MB per launch at 5.0 M ties = 18.6 B per tie (round 1 charged 64 B per tie here, overstating the step)
        "k_edges_resim": e * (16 + 4),
    }


def step_bytes(stats: dict) -> float:
    return float(sum(kernel_bytes(stats).values()))
