"""Deterministic random-init weights (re-export).  The generator itself lives in the product package
(`jyutvoice_b200/synthetic.py`: it only creates tensors and is what bench.py uses for its synthetic
weights); the oracle and the tests import it from here."""
from jyutvoice_b200.synthetic import *  # noqa: F401,F403
from jyutvoice_b200.synthetic import _bias_shape, _draw  # noqa: F401
