"""CPU (no GPU): the host control code + oracle kernels pass the reference's own regression
configurations and its check_solution against the reference's stored golden eigenvectors."""
import pytest

import lunda_cases as L


@pytest.mark.parametrize("name", sorted(L.LUNDA))
def test_lunda_hostcheck(name):
    L.run("hostcheck", name)
