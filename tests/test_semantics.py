"""CPU suite: the hand-derived quirk programs of tests/semantic_cases.py on the oracle."""
import pytest

import semantic_cases


@pytest.mark.parametrize("case", semantic_cases.ALL, ids=lambda c: c.__name__)
def test_semantic_case_on_oracle(case, oracle_mod):
    case(oracle_mod.OracleBatch)
