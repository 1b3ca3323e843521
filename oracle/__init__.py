"""CPU oracle (TEST INFRASTRUCTURE ONLY).

A dependency-free C++11 restatement of the reference's CPU tracker lives in ``ellc_oracle.cpp``; this package
is its ctypes binding.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it -- as the checker or the reported CPU baseline, never as the product
path.  Parity status: pinned on the reference's own code -- its unmodified tracking sources compile against the stand-in
OpenCV / Eigen / Boost headers of ``oracle/shim/`` (``ref_driver.cpp`` -> ``_ref/libellc_ref.so``, binding ``refbinding.py``)
and this oracle is bit-identical to that library at every iteration (``tests/test_reference_pin.py``); the third-party
arithmetic inside the stand-ins is pinned against cv2 / scipy fixtures under ``tests/golden``.
"""
from .binding import *  # noqa: F401,F403
