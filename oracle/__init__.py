"""CPU oracle for the robot-graph message-passing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(``multi-robot-perception-gnn-1_b200/``) imports this directory.  Allowed
importers: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` (as the checker / the timed CPU
baseline, never as the shipped path).

Parity status: the reference (xjh19971/multi-robot-perception-gnn-1) ships no
tests and no golden vectors for this path -> the oracle is pinned
against outputs of the *unmodified reference classes* (``dgl/model/models.py``:
``GCN``, ``edge_encoder``, ``edge_udf``, ``node_udf``, ``multi_view_dgl_model``'s
GNN slice) executed in the build container over ``oracle/dgl_shim`` (our
restatement of the DGL >= 0.5 semantics those classes rely on; real DGL is not
installable here).  ``tests/golden/make_golden.py`` is the generating script and
``tests/golden/*.npz`` the frozen vectors.  The DGL runtime itself is a
third-party dependency absent from /root/reference (version unpinned by the
reference; API implies DGL 0.5.x/0.6.x) - its batching / degree-bucketing
semantics are restated from its published behaviour, so for the graph-runtime
part parity is "pinned to the restated DGL semantics", not to a DGL binary.

``oracle/_ref/`` (git-ignored; written by ``oracle/build_ref.py`` from ``__graft_entry__.build()`` where /root/reference
is mounted) holds the reference's own ``dgl/model/{models,blocks}.py`` and ``dgl/{training,dataloader,dataloader_utils,
utils}.py`` as marshalled code objects - compiled in place, no reference source text in the repository.
``oracle/ref_runner.py`` executes them: the unmodified model classes are the timed ``--impl reference`` arm /
``cpu_baseline`` of bench.py and a second pin for the goldens; the unmodified ``train_dgl`` / ``test_dgl`` loops are run
over the compat packages by tests/test_compat.py.
"""
