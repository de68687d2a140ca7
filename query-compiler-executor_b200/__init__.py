"""B200-native execution engine for the per-query data-parallel path of
giorgosLiako/Query-Compiler-Executor (filter scan -> radix sort -> sort-merge
join -> row-id intermediates -> uint64 checksums).

Layout:
  csrc/   CUDA kernels (sm_100a) + the C-ABI implementation  -> libqce_b200.so
  src/    host-side operator layer in C, same entry points as the reference's
          src/*.h (filter.h, join.h, utilities.h, parsing.h, structs.h)
  main/   stdin->stdout driver (the reference's own main links unchanged too)
  engine.py   ctypes binding used by tests and bench.py

The directory name carries a hyphen; import it as `qce_b200` (shim at the repo
root).
"""
from .engine import Engine, EngineError, LIB_PATH, SYMBOLS, load_library  # noqa: F401
