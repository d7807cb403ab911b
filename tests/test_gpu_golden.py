"""The reference's known-answer vectors through the CUDA backend: device-resident arrays, host
(pinned) arrays staged inside the call, and the generic reference-ordered kernel forced on."""
import pytest

from golden_runner import CASE_IDS, CASES, run_case
from gpu_backend import GpuBackend

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
def test_golden_device(nc, case):
    run_case(case, GpuBackend(nc))


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
def test_golden_generic_kernel(nc, case):
    nc.lib().ncl_force_generic(1)
    try:
        run_case(case, GpuBackend(nc))
    finally:
        nc.lib().ncl_force_generic(0)


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
def test_golden_host_arrays(nc, case):
    run_case(case, GpuBackend(nc, where="host"))
