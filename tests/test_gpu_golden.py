"""GPU: the reference's own known answers (README.md, test/runtests.jl), run through the CUDA library
via the host mirror — the same checks that pin the oracle in test_oracle_golden.py."""
import numpy as np
import pytest

import known_answers as ka

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(fc):
    return ka.Api(
        chebpoints=fc.chebpoints,
        interp=lambda vals, lb, ub, tol: fc.chebinterp(vals, lb, ub, tol=tol),
        value=lambda c, x: c(x),
        grad=lambda c, x: fc.chebgradient(c, x),
        jac=lambda c, x: fc.chebjacobian(c, x),
        interp_v1=lambda vals, lb, ub: fc.chebinterp_v1(vals, lb, ub),
        coefs=lambda c: c.coefs,
        ArgumentError=fc.ArgumentError, DomainError=fc.DomainError, DimensionMismatch=fc.DimensionMismatch)


@pytest.mark.parametrize("check", ka.ALL_CHECKS, ids=lambda f: f.__name__)
def test_reference_known_answers(api, check):
    check(api)


@pytest.mark.parametrize("T", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("check", ka.TYPED_CHECKS, ids=lambda f: f.__name__)
def test_reference_testsets(api, check, T):
    check(api, T)
