"""The CPU oracle against golden vectors produced by the reference ITSELF:
the unmodified lib/fosphor/cl.c + fft.cl + display.cl run through NVIDIA's
OpenCL on the B200 box (tests/golden/make_golden.py, oracle/ref_build/).
This is what pins the oracle."""
import pytest

import golden_cases
import golden_check

CASES = ["cfg1", "sequence", "frame8", "kat_tones", "bh_range", "zeros_then_data", "einval"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    golden_check.check_case(name, golden_cases.OracleAdapter())
