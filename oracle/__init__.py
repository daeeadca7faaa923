"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference AP-VAST block engine.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and only as the checker or the reported CPU baseline.  The product path
(``ap_vast_unofficial_b200``) never imports this package and fails loudly when its CUDA
library is missing.
"""
