"""B200-native hot path of tduretz/FCFV_NME23 (FCFV Stokes: local coefficients, face-to-face assembly into device CSR,
residual evaluation) behind the C ABI of ``include/fcfv_b200.h``.

``api``        the reference's function names on ``mesh.FCFV_Mesh`` (Python mirror of ``julia/FCFV_B200.jl``)
``lib``        ctypes binding of ``libfcfv_b200.so`` (built in-tree from ``csrc/``; there is no CPU fallback)
``mesh``       the mesh struct, fixture readers, numbering helpers
``partition``  element partitioner for one-process-per-GPU runs
``solvers``    mirror of the consumer ``StokesSolvers!``: host factorisation (SciPy) around the GPU Powell-Hestenes lines
"""
__version__ = "0.1.0"
