"""B200-native training hot path for MMLRec (import as ``mmlrec_b200``).

Sub-modules: ``lib`` (ctypes binding of the C-ABI CUDA library), ``engine`` (static step
program: flat parameter store, workspace, kernel launch sequences, CUDA-graph capture),
``model`` (the reference's model classes / ``compile`` / ``fit`` / ``predict`` surface),
``utils`` (data pipeline surface), ``synthetic`` (BASELINE-shaped synthetic workloads).
Nothing here imports ``oracle/``; without the CUDA library the compute entry points raise.
"""
__version__ = "0.1.0"
