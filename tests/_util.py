import numpy as np


def assert_close_normwise(actual, expected, rtol, what=""):
    """|a - b| <= rtol * (|b| + max|b|): the tolerance form SURVEY.md section 7
    derives (elementwise relative error is ill-posed where a VACF crosses 0)."""
    actual, expected = np.asarray(actual), np.asarray(expected)
    assert actual.shape == expected.shape, f"{what}: shape {actual.shape} vs {expected.shape}"
    bound = rtol * (np.abs(expected) + np.abs(expected).max())
    err = np.abs(actual - expected)
    worst = np.argmax(err - bound)
    assert np.all(err <= bound), (
        f"{what}: max violation at {np.unravel_index(worst, err.shape)}: "
        f"|diff|={err.flat[worst]:.3e} bound={bound.flat[worst]:.3e}"
    )
