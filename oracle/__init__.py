"""CPU oracle for the next-POI training hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, what the reference's Theano graphs compute
for one ``model.train(...)`` / ``model.predict(...)`` call (reference files are
cited per function).  It exists so the CUDA engine can be checked against an
independent statement of the same arithmetic.

PINNED TO THE REFERENCE'S OWN MODEL CODE (Theano's arithmetic itself is not exercised): the reference
(tangrizzly/Point-of-Interest-Recommendation) ships no tests, golden vectors or recorded outputs, and its
arithmetic lives in Theano (un-pinned third-party dependency), which can neither be imported nor installed in
the build container.  The reference's model files do run under Python 3 unmodified once ``import theano``
resolves to ``oracle/theano_shim.py`` (a torch-backed evaluator of the API subset they use), so
``tests/golden/make_ref_golden.py`` executes the reference classes themselves and commits their outputs as
``tests/golden/ref_*.npz``; this package reproduces them to <= 3e-16 (``tests/test_golden.py``).  In addition:
(a) two independent statements of the same maths must agree (``oracle.models`` = torch autograd standing in for
``T.grad``; ``oracle.explicit`` = hand-derived backward in numpy), and (b) the analytic known answers listed in
``tests/test_oracle_known_answers.py``.

Nothing outside ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package.  The
product (``point-of-interest-recommendation_b200``) never does.
"""
