"""Oracle difficulty metrics vs values computed by the unmodified reference (metrics.npz),
including the reference's only known-answer material: the literal 15x15 maze."""
import numpy as np
import pytest

from conftest import load_golden
from oracle.metrics import kim_crawfis, mcclendon


def _rows():
    return [m for m in load_golden("metrics")[1] if not m["no_border"]]


def test_literal_maze_known_answers(golden_metrics):
    z, meta = golden_metrics
    m = meta[0]
    assert m["algo"] == "literal"
    # BASELINE.md section 2 / SURVEY.md section 8(c): reference outputs on testing_Mccledon.py:4-20
    assert m["difficulty"] == 9.950639302928026 and m["complexity"] == 5.681612603202764
    assert m["sol_len"] == 61 and m["L"] == 0.6288659793814433 and m["DE"] == 0.03278688524590164
    d = mcclendon(z["m0_grid"], m["start"], m["goal"], details=True)
    assert d["difficulty"] == pytest.approx(9.950639302928026, rel=1e-12)
    assert d["complexity"] == pytest.approx(5.681612603202764, rel=1e-12)
    assert d["n_hallways"] == 6 and d["n_branches"] == 3
    k = kim_crawfis(z["m0_grid"], m["start"], m["goal"])
    assert k["sol_len"] == 61 and k["L"] == 0.6288659793814433
    assert k["DE"] == 0.03278688524590164 and k["D"] == 0.03278688524590164


@pytest.mark.parametrize("m", _rows(), ids=lambda m: f"{m['algo']}-{m['shape']}-{m['id']}")
def test_metrics_match_reference(golden_metrics, m):
    z, _ = golden_metrics
    grid = z[f"m{m['id']}_grid"]
    d = mcclendon(grid, m["start"], m["goal"], details=True)
    assert d["n_hallways"] == m["n_hallways"]
    assert d["n_branches"] == m["n_branches"]
    assert d["hall_sum"] == pytest.approx(m["hall_sum"], rel=1e-11)
    assert d["difficulty"] == pytest.approx(m["difficulty"], rel=1e-11)
    assert d["complexity"] == pytest.approx(m["complexity"], rel=1e-11)
    k = kim_crawfis(grid, m["start"], m["goal"])
    assert k["sol_len"] == m["sol_len"]
    assert k["L"] == m["L"] and k["D"] == m["D"] and k["DE"] == m["DE"]   # exact: integer counts / same divisions


def test_no_border_difficulty_is_taken_on_the_bordered_maze(golden_metrics):
    """gen_maze_no_border (lib/maze_generation.py:48-56) evaluates difficulty before stripping."""
    z, meta = golden_metrics
    rows = [m for m in meta if m["no_border"]]
    assert rows
    for m in rows:
        bordered = np.pad(z[f"m{m['id']}_grid"], 1)
        d, _ = mcclendon(bordered, (m["start"][0] + 1, m["start"][1] + 1), (m["goal"][0] + 1, m["goal"][1] + 1))
        assert d == pytest.approx(m["difficulty"], rel=1e-11)


def _ext_rows():
    import json
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "metrics_ext.npz"))
    return json.loads(str(z["meta"]))


@pytest.mark.parametrize("row", _ext_rows(), ids=lambda r: f"m{r['id']}")
def test_extended_kim_crawfis_metrics_match_reference(golden_metrics, row):
    """density, T, J, CR, AC / FDE / BDE, L_DE, T_DE / D_sharp / L_sharp (metrics_calculator.py:18-69,
    175-244) as the unmodified MetricsCalculator computes them: bit-identical (same divisions, sums in
    the same row-major dead-end order)."""
    from oracle.metrics import kim_crawfis_extended
    z, meta = golden_metrics
    m = next(x for x in meta if x["id"] == row["id"])
    got = kim_crawfis_extended(z[f"m{m['id']}_grid"], m["start"], m["goal"])
    for k, v in row.items():
        if k != "id":
            assert got[k] == v, k


def _dead_end_fixpoint(grid, start, goal):
    """The parallel formulation maze_difficulty uses for the order-dependent part of calculate_DE
    (csrc/maze_metrics.cuh, step 4c), in plain Python: iterate `counted` from all-true until nothing changes."""
    from oracle.metrics import _Tree
    t = _Tree(grid, start, goal)
    sol_len = len(t.sol)
    chains = []
    for de in t.dead_ends_off_solution():
        path = t.path_to_start(de)
        for i in range(1, sol_len - 1):
            if t.on_sol[path[i]]:
                path = path[:i]
                break
        chains.append([p for p in path[1:-1] if t.nb[p] > 2])
    n, counted, rounds = len(chains), [True] * len(chains), 0
    while True:
        rounds += 1
        rec = {}
        for i in range(n):
            if counted[i] and chains[i]:
                rec[chains[i][0]] = min(rec.get(chains[i][0], n), i)
        new = [not any(rec.get(p, n) < i for p in chains[i]) for i in range(n)]
        if new == counted:
            return sum(counted), rounds
        counted = new


def test_dead_end_count_as_a_fixpoint_equals_the_sequential_rule():
    """The reference decides dead end by dead end, in row-major order, whether it counts (metrics_calculator.py:
    100-127).  The kernel solves the same rule as a fixpoint; both must agree on every maze, and the number of
    rounds stays small."""
    import random
    from oracle.generation import gen_maze
    rng = random.Random(5)
    worst = 0
    for k in range(45):
        S = rng.choice([11, 21, 41])
        algo = rng.choice(["r-prim", "dfs", "prim&kill"])
        start, goal, grid = gen_maze((S, S), algo, rng)
        got, rounds = _dead_end_fixpoint(grid, start, goal)
        assert got == kim_crawfis(grid, start, goal)["dead_end_count"], (S, algo)
        worst = max(worst, rounds)
    z, meta = load_golden("metrics")
    for m in meta:
        if not m["no_border"] and m["shape"] <= 61:
            got, rounds = _dead_end_fixpoint(z[f"m{m['id']}_grid"], m["start"], m["goal"])
            assert got / m["sol_len"] == pytest.approx(m["DE"], rel=1e-12), m["id"]
            worst = max(worst, rounds)
    assert worst <= 6
