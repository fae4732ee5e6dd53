"""CPU side of the README-table check: the committed reference samples (metric_table.npz, made by
the unmodified reference) reproduce the README's published means, which pins what the GPU
distribution test compares against."""
import numpy as np

from conftest import GOLDEN

README = {"r-prim": (71.90, 8.43, 0.04, 1.34, 0.33), "prim&kill": (99.08, 10.16, 0.14, 0.14, 0.07), "dfs": (106.41, 12.24, 0.47, 0.05, 0.03)}


def test_reference_samples_reproduce_the_readme_table():
    z = np.load(f"{GOLDEN}/metric_table.npz")
    for algo, published in README.items():
        a = z[algo]
        assert a.shape == (120, 5) and np.isfinite(a).all()
        for c, pub in enumerate(published):
            tol = 4.5 * a[:, c].std(ddof=1) / np.sqrt(len(a)) + 0.01 + 0.03 * abs(pub)
            assert abs(a[:, c].mean() - pub) <= tol, (algo, c, a[:, c].mean(), pub)
