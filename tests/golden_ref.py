"""Known-answer vectors copied from the reference's own tests (gp/gp_test.go);
printed to 6 digits there, compared at abs 1e-6 as the reference does."""
import numpy as np

# TestProduce, gp/gp_test.go:23-120: (name, noise std, theta, x, y, z, mu, sigma); kernel.Normal, NDim 1
PRODUCE = [
    ("prior", 0.0, [1.0], [], [], [[0.0]], [0.0], [1.0]),
    ("self", 0.0, [1.0], [[0.0]], [1.0], [[0.0]], [1.0], [0.0]),
    ("next", 0.0, [1.0], [[0.0]], [0.0], [[1.0]], [0.0], [0.795060]),
    ("two selves", 0.0, [1.0], [[0.0], [1.0]], [1.0, -1.0], [[0.0], [1.0]], [1.0, -1.0], [0.0, 0.0]),
    ("inter", 0.0, [1.0], [[0.0], [1.0]], [1.0, -1.0], [[0.5]], [0.0], [0.174518]),
    ("extra", 0.0, [1.0], [[0.0], [1.0]], [1.0, -1.0], [[-2.0], [3.0]], [0.315720, -0.315720], [0.986770, 0.986770]),
    ("noise", 0.1, [1.0], [[0.0], [1.0]], [1.0, -1.0], [[-2.0], [3.0]], [0.307895, -0.307895], [0.987037, 0.987037]),
]

# TestElementalModel, gp/gp_test.go:180-229: (name, noise, x, ll); noise is a std or "uniform"
ELEMENTAL = [
    ("prior", 0.0, [0.0], 0.0),
    ("single", 0.0, [0.0, 0.0, 1.0], -1.418939),
    ("nonoise", 0.0, [0.0, 0.0, 1.0, 1.0, 0.0], -2.399528),
    ("withnoise", 0.1, [1.0, -2.0, -1.0, 1.0, 0.0], -4.321055),
    ("uninoise", "uniform", [1.0, 1.0, -1.0, -1.0, 1.0, 0.0], -4.018110),
]
DX, EPS = 1e-8, 1e-4  # gp/gp_test.go:168-171


def arr(v):
    return np.array(v, dtype=np.float64)
