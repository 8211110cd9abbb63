"""The paper's parameter-count table (reference demo/figures.py:236-293) against counts produced by the unmodified reference
(tests/golden/params_kat.json, tests/golden/make_golden.py params): same weights (numpy seed 0), same key seed."""
import json
import os

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.slow]
HERE = os.path.dirname(os.path.abspath(__file__))


def test_parameter_table_matches_the_reference():
    import bench
    from keynet_b200 import report
    gold = dict((k, v) for (k, v) in json.load(open(os.path.join(HERE, 'golden', 'params_kat.json'))))
    rows = dict(report.parameter_table(('lenet',), init=lambda n: bench.numpy_weights(n, 0), verbose=False))
    rows.update(dict(report.parameter_table(('allconvnet',), tiles=[8], init=lambda n: bench.numpy_weights(n, 0), verbose=False)))
    for (k, v) in gold.items():
        assert rows[k] == v, (k, rows[k], v)
