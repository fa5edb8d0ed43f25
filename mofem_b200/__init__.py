"""mofem_b200 -- B200-native stiffness assembly + PCG solve for overlapping FE meshes (AMORE).

Drop-in for the hot path of junbinhuang/MOFEM: ``linearSolvers/AMORE.py`` and
``linearSolvers/traditionalElement.py`` entry points, backed by hand-written
sm_100a CUDA kernels behind a C-ABI shared library (``include/mofem_b200.h``).
There is no CPU fallback: every compute entry point raises if the CUDA library
or a GPU is missing.
"""
__version__ = "0.1.0"
