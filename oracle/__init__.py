"""CPU oracle for the ProtoASNet prototype-head / push hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``protoasnet_b200/`` may import this package.
Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs -- and there only as the checker or the timed CPU baseline,
never as the product path.

Parity pinning: the reference ships no tests and no golden vectors.  The
oracle is therefore pinned against *live outputs of the reference classes themselves*,
generated in the build container by ``oracle/gen_golden.py`` (which imports
``/root/reference/src/models/{Video_XProtoNet,XProtoNet}.py`` and runs the unmodified
``src/utils/push_abs_revision.py::push_prototypes`` on CPU behind import shims) and
committed under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every oracle
function against those fixtures.
"""
