"""CPU oracle of the DGOD detection-head hot path — TEST INFRASTRUCTURE ONLY.

`oracle.cpu` wraps `dgod_oracle.c` (a scalar C restatement of the torchvision CPU algorithms the
reference calls) with a numpy interface.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import this package; the product
package `dgod_b200` never does (tests/test_boundary.py enforces it).
"""
