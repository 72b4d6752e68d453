"""GPU: the product passes the reference's own regression configurations (tests/tests/test_001
.. test_005 on LUNDA.mtx) and check_solution against the reference's stored golden vectors."""
import pytest

import lunda_cases as L

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(L.LUNDA))
def test_lunda_product(name):
    r = L.run("product", name)
    assert r["launches"] > 0
