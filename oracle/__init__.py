"""CPU oracle (TEST INFRASTRUCTURE ONLY).

A dependency-free C++11 restatement of the reference's CPU tracker lives in ``ellc_oracle.cpp``; this package
is its ctypes binding.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it -- as the checker or the reported CPU baseline, never as the product
path.  Parity status: unpinned by the reference (it ships no tests/golden vectors and cannot be built here);
pinned against cv2 / scipy fixtures under ``tests/golden`` instead.
"""
from .binding import *  # noqa: F401,F403
