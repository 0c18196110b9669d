"""CPU oracle for the next-POI training hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, what the reference's Theano graphs compute
for one ``model.train(...)`` / ``model.predict(...)`` call (reference files are
cited per function).  It exists so the CUDA engine can be checked against an
independent statement of the same arithmetic.

PARITY UNPINNED: the reference (tangrizzly/Point-of-Interest-Recommendation)
ships no tests, golden vectors or recorded outputs, and its arithmetic lives in
Theano (un-pinned third-party dependency, Python 2 only), which can neither be
imported nor installed in the build container.  The oracle is therefore pinned
only by (a) two independent statements of the same maths that must agree
(``oracle.models`` = torch autograd standing in for ``T.grad``;
``oracle.explicit`` = hand-derived backward in numpy; ``oracle/c`` = plain C),
and (b) the analytic known answers listed in ``tests/test_oracle_known_answers.py``.

Nothing outside ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package.  The
product (``point-of-interest-recommendation_b200``) never does.
"""
