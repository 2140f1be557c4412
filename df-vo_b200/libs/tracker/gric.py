"""GRIC model-selection scores with the reference's function names (libs/tracker/gric.py:14-132);
the arithmetic lives in b200/hostmath.py (O(N), vectorised) and, for the essential-matrix candidates of
the RANSAC repeats, in the k_finalize CUDA kernel."""
from b200.hostmath import calc_gric as calc_GRIC                                   # noqa: F401
from b200.hostmath import fundamental_residual as compute_fundamental_residual       # noqa: F401
from b200.hostmath import homography_residual as compute_homography_residual         # noqa: F401
