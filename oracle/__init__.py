"""CPU oracle for the post-encoder geolocation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker or as the timed CPU baseline -- never behind the product API.  The
product (``geoguessr_ai_b200``) fails loudly when its CUDA library is missing;
it never routes through this package.

The oracle is a restatement, in eager CPU PyTorch (the reference's own
arithmetic: the reference has no native code, every FLOP on this path is an
ATen call), of

* ``models/super_guessr.py:336-395``  (fusion, head, softmax, argmax, top-k,
  haversine label-smoothed CE)                       -> ``super_guessr_oracle``
* ``models/utils.py:20-57``           (``smooth_labels``, ``haversine_matrix``)
* ``models/proto_refiner.py:129-237,364-389`` and
  ``preprocessing/geo_utils.py:39-54`` (refiner forward) -> ``proto_refiner_oracle``
* ``models/proto_refiner.py:461-517`` (cluster prototypes from member embeddings; the
  encoder call replaced by stored embeddings)         -> ``proto_builder_oracle``

Pinning: the reference's own tests hold no vector for this path (its only
collected test is ``assert True``), so the oracle is pinned against the
reference ITSELF, imported and executed in the build container by
``oracle/make_golden.py`` (which needs ``/root/reference``); the resulting
vectors are committed under ``tests/golden/`` and ``tests/test_oracle.py``
re-checks the oracle against them on every run.  The same holds for the two
host-side formats next to the path: ``make_golden_sqlite.py`` executes the
reference's panorama grouping (``training/load_sqlite_dataset.py:104-150``) and
``make_golden_protos.py`` its ``ProtoDataManager`` (``models/utils.py:98-181``).
"""
