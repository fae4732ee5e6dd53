"""Difficulty metrics oracle (test infrastructure; see oracle/__init__.py).

McClendon complexity / difficulty  (lib/maze_difficulty_evaluation/maze_complexity_evaluation.py)
Kim & Crawfis L / D / DE           (lib/maze_difficulty_evaluation/metrics_calculator.py)

This is a restatement on the maze's spanning tree, not a port of the reference's A* + networkx
code: a BFS from `start` gives parent pointers; the reference's graph G is exactly the maze tree
compressed onto its "nodes" (start, dead ends, corners, junctions, crossings), so every quantity
is a sum over tree edges.  The reference's order-dependent quirks are honoured explicitly (see
`_first_child`).  Pinned against reference outputs in tests/golden/metrics.npz.
"""
from __future__ import annotations

import math

import numpy as np

from .grid import bfs_dist

_N4 = ((-1, 0), (1, 0), (0, -1), (0, 1))


class _Tree:
    """Spanning tree of a bordered perfect maze rooted at `start` (block resolution)."""

    def __init__(self, grid, start, goal):
        self.g = np.asarray(grid)
        self.H, self.W = self.g.shape
        self.start = (int(start[0]), int(start[1]))
        self.goal = (int(goal[0]), int(goal[1]))
        self.ds = bfs_dist(self.g, self.start)
        g = (self.g != 0).astype(np.int32)
        nb = np.zeros_like(g)
        nb[1:-1, 1:-1] = g[:-2, 1:-1] + g[2:, 1:-1] + g[1:-1, :-2] + g[1:-1, 2:]
        self.nb = nb * g
        # solution blocks, start -> goal
        sol = [self.goal]
        while sol[-1] != self.start:
            sol.append(self.parent(sol[-1]))
        sol.reverse()
        self.sol = sol
        self.on_sol = np.zeros(self.g.shape, dtype=bool)
        for b in sol:
            self.on_sol[b] = True

    def parent(self, b):
        r, c = b
        d = self.ds[r, c]
        for dr, dc in _N4:
            n = (r + dr, c + dc)
            if self.g[n] != 0 and self.ds[n] == d - 1:
                return n
        raise ValueError("no parent (not a tree rooted at start?)")

    def path_to_start(self, b):
        out = [b]
        while out[-1] != self.start:
            out.append(self.parent(out[-1]))
        return out

    def dead_ends_off_solution(self):
        """Row-major list of cells with value 1, one open neighbour, not on the solution
        (maze_complexity_evaluation.py:152-166, metrics_calculator.py:129-138)."""
        out = []
        for r in range(1, self.H - 1):
            for c in range(1, self.W - 1):
                if self.g[r, c] == 1 and self.nb[r, c] == 1 and not self.on_sol[r, c]:
                    out.append((r, c))
        return out

    def is_node(self, b, prev_b, next_b):
        """decompose_in_turns (maze_complexity_evaluation.py:125-136): interior path block is a
        node if the path turns there or it has more than two open neighbours."""
        turn = prev_b[0] != next_b[0] and prev_b[1] != next_b[1]
        return turn or self.nb[b] > 2


def mcclendon(grid, start, goal, details=False):
    """-> (difficulty, complexity)  [difficulty_of_maze :319-329, complexity_of_maze :310-317]."""
    t = _Tree(grid, start, goal)
    # ---- compressed tree: node set and parent edges -------------------------------------------
    # nodes: path endpoints (start, goal, off-solution dead ends) + turns + >2-neighbour blocks
    parent_node, dist_to_parent, kids = {}, {}, {}
    sol_nodes = []

    def add_chain(path_from_leaf_to_start):
        """Register the nodes of one leaf->start path; returns node list in path order."""
        p = path_from_leaf_to_start
        nodes = [p[0]]
        for i in range(1, len(p) - 1):
            if t.is_node(p[i], p[i - 1], p[i + 1]):
                nodes.append(p[i])
        nodes.append(p[-1])
        pos = {b: i for i, b in enumerate(p)}
        for a, b in zip(nodes[:-1], nodes[1:]):       # a below, b above (towards start)
            if a not in parent_node:
                parent_node[a] = b
                dist_to_parent[a] = pos[b] - pos[a] - 1   # blocks strictly between (:176-184)
                kids.setdefault(b, []).append(a)          # insertion order == networkx adjacency order
        return nodes

    sol_rev = list(reversed(t.sol))                       # goal -> start
    sol_nodes = list(reversed(add_chain(sol_rev)))        # start -> goal order (:62-66)
    for de in t.dead_ends_off_solution():                 # row-major (:71-80)
        add_chain(t.path_to_start(de))
    nodes = set(parent_node) | {t.start}
    sol_set = set(sol_nodes)
    junction = {n for n in nodes if t.nb[n] == 3}         # exactly three open neighbours (:138-150)

    def edge_terms(d):
        return d, 1.0 / (2 * d)

    # ---- hallway 0: the solution chain --------------------------------------------------------
    D0 = S0 = 0.0
    for n in sol_nodes[1:]:
        d, s = edge_terms(dist_to_parent[n])
        D0 += d; S0 += s
    c_solution = D0 * S0

    # ---- hallways >= 1: components of G - junctions - solution nodes (:186-221) --------------
    plain = [n for n in nodes if n not in junction and n not in sol_set]
    comp_root = {}

    for n in plain:
        # walk up while the parent is plain
        r = n
        while True:
            p = parent_node[r]
            if p in junction or p in sol_set:
                break
            r = p
        comp_root[n] = r
    comps = {}
    for n in plain:
        comps.setdefault(comp_root[n], []).append(n)

    def first_child(p):
        """Child of p reached first in networkx adjacency order = the one on the path of the
        row-major-first dead end below p (dead-end chains are inserted in row-major order)."""
        return kids[p][0]

    sums = {r: [0.0, 0.0] for r in comps}
    for n in nodes:
        if n == t.start:
            continue
        p = parent_node[n]
        d, s = edge_terms(dist_to_parent[n])
        if n in comp_root:
            # plain child: edge is in its component's hallway if the parent is plain or a junction
            if p in comp_root or p in junction:
                sums[comp_root[n]][0] += d; sums[comp_root[n]][1] += s
        elif n in junction and n not in sol_set and p in comp_root:
            # junction child of a plain node: reached by p's neighbour scan unless the scan broke
            # on p's parent being a junction on the solution (:209-214) before reaching this child
            pp = parent_node[p]
            broke = pp in junction and pp in sol_set
            if not broke or first_child(p) == n:
                sums[comp_root[p]][0] += d; sums[comp_root[p]][1] += s

    # ---- branches: components of G - (solution nodes that are not junctions) (:223-259) ------
    sol_index = {n: i for i, n in enumerate(sol_nodes)}

    def branch_key(r):
        below, x = r, parent_node[r]
        while x not in sol_set:
            below, x = x, parent_node[x]
        if x in junction:
            i = sol_index[x]
            while i > 0 and sol_nodes[i - 1] in junction:
                i -= 1
            return ("run", sol_nodes[i])
        return ("sub", below)

    branch_sum = {}
    for r, (Dh, Sh) in sums.items():
        k = branch_key(r)
        branch_sum[k] = branch_sum.get(k, 0.0) + Dh * Sh

    total = c_solution
    prod = c_solution
    for v in branch_sum.values():
        total += v
        prod *= v + 1
    difficulty, complexity = math.log(prod), math.log(total)
    if details:
        # branch components also exist without hallways (e.g. junction runs whose subtrees are all junctions)
        keys = set(branch_sum)
        for n in nodes:
            if n in sol_set and n in junction:
                i = sol_index[n]
                while i > 0 and sol_nodes[i - 1] in junction:
                    i -= 1
                keys.add(("run", sol_nodes[i]))
            elif n not in sol_set and parent_node[n] in sol_set and parent_node[n] not in junction:
                keys.add(("sub", n))
        return dict(difficulty=difficulty, complexity=complexity, n_hallways=1 + len(comps), n_branches=1 + len(keys),
                    hall_sum=c_solution + sum(D * S for D, S in sums.values()))
    return difficulty, complexity


def kim_crawfis(grid, start, goal):
    """-> dict(L, D, DE, sol_len)  [calculate_L :22-26, calculate_D :71-85, calculate_DE :87-127]."""
    t = _Tree(grid, start, goal)
    sol_len = len(t.sol)
    ce = (t.H - 1) * ((t.W - 1) // 2) - 1                       # :16
    L = sol_len / ce
    D = sum(1 for b in t.sol if t.nb[b] > 2) / sol_len
    # dead ends, row-major; a dead end counts unless its (cut) path to the solution holds a
    # decision point recorded by an earlier counted dead end (:107-116)
    recorded = set()
    alcoves = forward = backward = 0
    gr, gc = t.goal
    for de in t.dead_ends_off_solution():
        path = t.path_to_start(de)
        # calculate_path :140-151 -- cut at the first solution block, but only if it sits at an
        # index <= sol_len - 2 (the loop bound is len(solution), not len(de_path))
        for i in range(1, sol_len - 1):
            if t.on_sol[path[i]]:
                path = path[:i]
                break
        if set(path) & recorded:
            continue
        for k in range(1, len(path) - 1):
            if t.nb[path[k]] > 2:
                recorded.add(path[k])
                break
        # type_of_DE :153-173: alcove unless the path has an interior decision block or a turn
        interior = range(1, len(path) - 1)
        has_turn = any(path[i - 1][0] != path[i + 1][0] and path[i - 1][1] != path[i + 1][1] for i in interior)
        flag = len(path) >= 3 and (has_turn or any(t.nb[path[k]] > 2 for k in interior))
        if flag:
            diff = (abs(path[-1][0] - gr) + abs(path[-1][1] - gc)) - (abs(path[0][0] - gr) + abs(path[0][1] - gc))
            if diff > 0:
                forward += 1
            else:
                backward += 1
        else:
            alcoves += 1
    DE = alcoves / sol_len + forward / sol_len + backward / sol_len      # :97-98, summed in this order
    return dict(L=L, D=D, DE=DE, sol_len=sol_len, dead_end_count=alcoves + forward + backward,
                alcoves=alcoves, forward=forward, backward=backward)


DE_TYPES = ("AC", "FDE", "BDE")


def kim_crawfis_extended(grid, start, goal):
    """The MetricsCalculator methods the reference defines but never calls
    (metrics_calculator.py:18-69,175-244), restated on the spanning tree:

      density :18-20   open blocks / (H * W)
      T :28-37  J :39-53  CR :55-69   turns / 3-neighbour / 4-neighbour blocks of the solution over len(sol)
      AC, FDE, BDE :100-127           the three terms of DE
      L_DE :224-241                   sum over ALL off-solution dead ends of len(de_path) / CE
      T_DE(type) :175-185             sum over dead ends of that type of (turns(de_path) / len(sol)) / len(de_path)
      D_sharp(type) :187-197          same with the > 2-neighbour blocks of de_path (end points included)
      L_sharp(type) :199-222          sum over dead ends of that type of len(de_path) / CE

    `find_decision` (:243-255) iterates `range(1, len(path) - 1, -1)`, which is empty for every path
    of two or more blocks, so it always returns None: L_DE and L_sharp never shorten a path and never
    record a decision point.  Sums run in row-major dead-end order with the reference's operations,
    so the values are bit-identical, not just close."""
    t = _Tree(grid, start, goal)
    sol, sol_len = t.sol, len(t.sol)
    ce = (t.H - 1) * ((t.W - 1) // 2) - 1
    gr, gc = t.goal

    def turns(path):
        return sum(1 for i in range(1, len(path) - 1)
                   if path[i - 1][0] != path[i + 1][0] and path[i - 1][1] != path[i + 1][1])

    out = dict(density=int((t.g != 0).sum()) / (t.H * t.W), T=turns(sol) / sol_len,
               J=sum(1 for b in sol if t.nb[b] == 3) / sol_len, CR=sum(1 for b in sol if t.nb[b] == 4) / sol_len)
    base = kim_crawfis(grid, start, goal)
    out.update(AC=base["alcoves"] / sol_len, FDE=base["forward"] / sol_len, BDE=base["backward"] / sol_len)
    L_DE = 0
    T_DE, D_sharp, L_sharp = [0, 0, 0], [0, 0, 0], [0, 0, 0]
    for de in t.dead_ends_off_solution():
        path = t.path_to_start(de)
        for i in range(1, sol_len - 1):            # calculate_path :140-151
            if t.on_sol[path[i]]:
                path = path[:i]
                break
        n_turns = turns(path)
        interior_dp = any(t.nb[path[k]] > 2 for k in range(1, len(path) - 1))
        if len(path) >= 3 and (interior_dp or n_turns > 0):           # type_of_DE :153-173
            diff = (abs(path[-1][0] - gr) + abs(path[-1][1] - gc)) - (abs(path[0][0] - gr) + abs(path[0][1] - gc))
            k = 1 if diff > 0 else 2
        else:
            k = 0
        L_DE += len(path) / ce
        L_sharp[k] += len(path) / ce
        T_DE[k] += (n_turns / sol_len) / len(path)
        D_sharp[k] += (sum(1 for b in path if t.nb[b] > 2) / sol_len) / len(path)
    out.update(L_DE=L_DE, T_DE=T_DE, D_sharp=D_sharp, L_sharp=L_sharp)
    return out
